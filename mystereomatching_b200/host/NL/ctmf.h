// ctmf.h -- same C signature as the reference's NL/ctmf.h:8; the body runs the sm_100a median kernel.
#ifndef CTMF_H
#define CTMF_H
#ifdef __cplusplus
extern "C" {
#endif
/* (2r+1)^2 median of an 8-bit image with `channels` interleaved channels, edge-replicated border -- what
 * NL/ctmf.c:378-433 computes.  memsize (the CPU cache-blocking hint) is ignored.  r in {1,2,3}. */
void ctmf(const unsigned char* src, unsigned char* dst, int width, int height, int src_step_row, int dst_step_row, int r,
          int channels, unsigned long memsize);
#ifdef __cplusplus
}
#endif
#endif
