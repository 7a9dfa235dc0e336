"""What the scratch file system itself delivers (the denominator of the file -> file figure): T threads pwrite / pread
slices of F files of G GB each, new files (page allocation included) and existing ones.
  python tools/fs_peak.py [dir] [GB per file] """
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np

d = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
G = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
SL = 4 << 20
buf = np.full(SL, 65, np.uint8)
n = int(G * 1e9) // SL


def run(files, threads, op, fresh):
    paths = [os.path.join(d, "bk_fs_peak_%d" % i) for i in range(files)]
    if fresh:
        for p in paths:
            if os.path.exists(p):
                os.remove(p)
    fds = [os.open(p, os.O_RDWR | os.O_CREAT, 0o644) for p in paths]
    rb = [np.empty(SL, np.uint8) for _ in range(threads)]

    def task(i):
        f, s = i % files, i // files
        if op == "w":
            os.pwrite(fds[f], buf, s * SL)
        else:
            os.preadv(fds[f], [rb[i % threads]], s * SL)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(task, range(n * files)))
    dt = time.perf_counter() - t0
    for fd in fds:
        os.close(fd)
    return n * files * SL / dt / 1e9


def run_mmap(files, threads):
    """new files, written through a shared mapping (page faults instead of write(2): no inode lock)"""
    import mmap
    paths = [os.path.join(d, "bk_fs_peak_%d" % i) for i in range(files)]
    for p in paths:
        if os.path.exists(p):
            os.remove(p)
    fds = [os.open(p, os.O_RDWR | os.O_CREAT, 0o644) for p in paths]
    maps = []
    for fd in fds:
        os.ftruncate(fd, n * SL)
        maps.append(np.frombuffer(mmap.mmap(fd, n * SL), np.uint8))

    def task(i):
        f, s = i % files, i // files
        np.copyto(maps[f][s * SL:(s + 1) * SL], buf)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(task, range(n * files)))
    dt = time.perf_counter() - t0
    del maps
    for fd in fds:
        os.close(fd)
    return n * files * SL / dt / 1e9


for files in (1, 2):
    for threads in (1, 4, 8, 16):
        print("%d file(s), %2d threads: write new through mmap %5.1f GB/s" % (files, threads, run_mmap(files, threads)), flush=True)
for files in ():
    for threads in (1, 2, 4, 8, 16):
        w = run(files, threads, "w", True)
        ow = run(files, threads, "w", False)
        r = run(files, threads, "r", False)
        print("%d file(s), %2d threads: write new %5.1f GB/s, overwrite %5.1f GB/s, read %5.1f GB/s" % (files, threads, w, ow, r), flush=True)
for i in range(2):
    p = os.path.join(d, "bk_fs_peak_%d" % i)
    if os.path.exists(p):
        os.remove(p)
