/* C11 driver of the drop-in boundary, compiled with gcc against include/barkit_b200.h (no C++, no Python):
 *   abi_driver sizes                      -> struct sizes as the C compiler sees them
 *   abi_driver nogpu                      -> pattern front end + the loud failure of bk_ctx_create without a device
 *   abi_driver extract IN.fq OUT.fq PATTERN K  -> bk_tokenize, bk_compile, bk_ctx_create, bk_submit, bk_wait (single end)
 * tests/test_abi_c.py builds and runs it; the extract mode's output is compared with the oracle there. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "barkit_b200.h"

static int fail(const char* what, int rc, const char* msg) {
  fprintf(stderr, "abi_driver: %s failed (%d): %s\n", what, rc, msg ? msg : "");
  return 1;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "sizes")) {
    printf("bk_mate_in %zu\nbk_result %zu\nbk_run_args %zu\nbk_program_info %zu\nbk_syn_config %zu\n", sizeof(bk_mate_in),
           sizeof(bk_result), sizeof(bk_run_args), sizeof(bk_program_info), sizeof(bk_syn_config));
    return 0;
  }
  if (argc >= 2 && !strcmp(argv[1], "nogpu")) {
    char buf[256], err[512];
    int st = 0;
    if (bk_expand("^atgc(?<UMI>[ATGCN]{12})", 1, buf, sizeof buf) != BK_OK) return fail("bk_expand", -1, buf);
    printf("expand %s\n", buf);  /* pattern.rs:88-91 */
    bk_program* p = bk_compile("^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})", 1, &st, err, sizeof err);
    if (!p) return fail("bk_compile", st, err);
    printf("caps %d %s %s\n", bk_program_ncaps(p), bk_program_cap_name(p, 0), bk_program_cap_name(p, 1));
    if (bk_compile("^(?P<XX>[ATGCN]{4})", 1, &st, err, sizeof err) != NULL || st != BK_E_PATTERN) return fail("bad name", st, err);
    printf("error %s\n", err);  /* error.rs:13-14 */
    bk_ctx* ctx = bk_ctx_create(0, p, NULL, 0, &st, err, sizeof err);
    if (ctx) { printf("ctx ok\n"); bk_ctx_destroy(ctx); }
    else printf("ctx %d %s\n", st, err);
    bk_program_destroy(p);
    return 0;
  }
  if (argc == 6 && !strcmp(argv[1], "extract")) {
    FILE* f = fopen(argv[2], "rb");
    if (!f) return fail("fopen", 0, argv[2]);
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* text = (uint8_t*)bk_host_alloc((size_t)len + 64);
    if (!text) return fail("bk_host_alloc", BK_E_CUDA, "pinned allocation");
    memset(text, 0, (size_t)len + 64);
    if (fread(text, 1, (size_t)len, f) != (size_t)len) return fail("fread", 0, argv[2]);
    fclose(f);
    uint32_t cap = (uint32_t)(len / 4 + 2), n = 0, maxrec = 0;
    uint64_t consumed = 0;
    uint32_t* off = (uint32_t*)malloc(((size_t)cap + 1) * 4);
    int rc = bk_tokenize(text, (uint64_t)len, 1, off, cap, &n, &consumed, &maxrec);
    if (rc != BK_OK || consumed != (uint64_t)len) return fail("bk_tokenize", rc, "malformed FASTQ");
    char err[512];
    int st = 0;
    bk_program* p = bk_compile(argv[4], (size_t)atoi(argv[5]), &st, err, sizeof err);
    if (!p) return fail("bk_compile", st, err);
    bk_ctx* ctx = bk_ctx_create(0, p, NULL, 0, &st, err, sizeof err);
    if (!ctx) return fail("bk_ctx_create", st, err);
    bk_mate_in in;
    memset(&in, 0, sizeof in);
    in.text = text;
    in.rec_off = off;
    in.text_len = off[n];
    in.max_rec_bytes = maxrec ? maxrec : 1;
    uint64_t ocap = bk_out_bound(ctx, 1, n, in.text_len) + 16;
    uint8_t* out = (uint8_t*)bk_host_alloc(ocap);
    bk_result res;
    memset(&res, 0xEE, sizeof res); /* every field must be written by bk_wait */
    if ((rc = bk_submit(ctx, 0, n, &in, NULL)) != BK_OK) return fail("bk_submit", rc, bk_last_error(ctx));
    if (bk_submit(ctx, 0, n, &in, NULL) == BK_OK) return fail("second bk_submit on a busy slot", 0, "accepted");
    if ((rc = bk_wait(ctx, 0, out, ocap, NULL, 0, &res)) != BK_OK) return fail("bk_wait", rc, bk_last_error(ctx));
    /* a too-small output buffer is an error that leaves the slot reusable */
    if ((rc = bk_submit(ctx, 1, n, &in, NULL)) != BK_OK) return fail("bk_submit(1)", rc, bk_last_error(ctx));
    bk_result res2;
    rc = bk_wait(ctx, 1, out, res.out1_len ? res.out1_len - 1 : 0, NULL, 0, &res2);
    if (res.out1_len && rc != BK_E_CAPACITY) return fail("bk_wait with a short buffer", rc, "expected BK_E_CAPACITY");
    if ((rc = bk_submit(ctx, 1, n, &in, NULL)) != BK_OK) return fail("bk_submit(1) after the error", rc, bk_last_error(ctx));
    if ((rc = bk_wait(ctx, 1, out, ocap, NULL, 0, &res2)) != BK_OK) return fail("bk_wait(1)", rc, bk_last_error(ctx));
    if (res2.out1_len != res.out1_len || res2.n_out != res.n_out) return fail("second run differs", 0, "");
    f = fopen(argv[3], "wb");
    if (!f || fwrite(out, 1, res.out1_len, f) != res.out1_len) return fail("fwrite", 0, argv[3]);
    fclose(f);
    printf("n_in %llu n_out %llu out1_len %llu launches %u h2d %llu d2h %llu\n", (unsigned long long)res.n_in,
           (unsigned long long)res.n_out, (unsigned long long)res.out1_len, res.launches,
           (unsigned long long)res.h2d_bytes, (unsigned long long)res.d2h_bytes);
    bk_host_free(out);
    bk_host_free(text);
    free(off);
    bk_ctx_destroy(ctx);
    bk_program_destroy(p);
    return 0;
  }
  fprintf(stderr, "usage: abi_driver sizes | nogpu | extract IN.fq OUT.fq PATTERN K\n");
  return 2;
}
