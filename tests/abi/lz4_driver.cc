// Test driver for barkit_b200/csrc/bk_lz4.h (compiled by tests/test_lz4_codec.py with g++):
//   lz4_driver c < text > frame     one frame, 1 MiB independent blocks (what `barkit extract --lz4` writes)
//   lz4_driver d < frame > text     the streaming decoder, fed 64 KiB at a time, read in odd-sized pieces
#include <stdio.h>
#include <unistd.h>

#include "../../barkit_b200/csrc/bk_lz4.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  if (argv[1][0] == 'c') {
    std::vector<uint8_t> in, out;
    uint8_t buf[1 << 16];
    ssize_t r;
    while ((r = read(0, buf, sizeof buf)) > 0) in.insert(in.end(), buf, buf + r);
    bklz4::frame_header(&out);
    for (size_t b = 0; b < in.size(); b += bklz4::kBlock) bklz4::frame_block(in.data() + b, std::min(bklz4::kBlock, in.size() - b), &out);
    for (int i = 0; i < 4; ++i) out.push_back(0);
    fwrite(out.data(), 1, out.size(), stdout);
    return 0;
  }
  bklz4::Decoder dec;
  std::string err;
  std::vector<uint8_t> piece(100003);
  for (;;) {
    const long r = dec.read(piece.data(), piece.size(), [](uint8_t* b, size_t n) { return (long)read(0, b, std::min<size_t>(n, 1 << 16)); }, &err);
    if (r < 0) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    fwrite(piece.data(), 1, (size_t)r, stdout);
    if ((size_t)r < piece.size()) break;
  }
  return 0;
}
