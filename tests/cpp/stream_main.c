/* A plain-C caller of the frame stream (include/sm_b200.h, sm_stream_*): what a maintainer's loop over stereo pairs
 * (main_.cpp:138-166, one StereoMatching per pair) becomes with this library.
 *   stream_main <in.bin> <out.bin> H W D paths n_frames n_devices
 * in.bin  = n_frames x { bgrL[H*W*3], bgrR[H*W*3], grayL[H*W], grayR[H*W] };  out.bin = n_frames x dispL[H*W] (int16).
 * Frames are submitted in order (frame i -> device i mod n_devices) and waited for in REVERSE order. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sm_b200.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc__ = (call);                                                     \
    if (rc__ != SM_OK) {                                                   \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc__, sm_last_error());    \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage: %s in out H W D paths n_frames n_devices\n", argv[0]); return 2; }
  const int H = atoi(argv[3]), W = atoi(argv[4]), D = atoi(argv[5]), paths = atoi(argv[6]), n = atoi(argv[7]);
  int ndev = atoi(argv[8]);
  if (ndev > sm_device_count()) ndev = sm_device_count();
  if (ndev < 1) { fprintf(stderr, "no CUDA device: this library has no CPU fallback\n"); return 3; }
  const size_t npix = (size_t)H * W, frameB = npix * 8;
  unsigned char* in = (unsigned char*)malloc(frameB * n);          /* pageable on purpose: goes through the staging */
  short* out = (short*)malloc(npix * 2 * n);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(in, 1, frameB * n, f) != frameB * n) { fprintf(stderr, "short read\n"); return 4; }
  fclose(f);
  sm_params p;
  sm_params_default(&p, D - 1);
  p.sgm_paths = paths;
  int devices[64];
  for (int i = 0; i < ndev; i++) devices[i] = i;
  sm_stream* s = NULL;
  CHECK(sm_stream_create(devices, ndev, H, W, &p, 3, &s));
  long long* tickets = (long long*)malloc(sizeof(long long) * n);
  for (int i = 0; i < n; i++) {
    const unsigned char* fr = in + frameB * i;
    CHECK(sm_stream_submit(s, fr, fr + npix * 3, fr + npix * 6, fr + npix * 7, out + npix * i, &tickets[i]));
  }
  for (int i = n - 1; i >= 0; i--) CHECK(sm_stream_wait(s, tickets[i]));
  CHECK(sm_stream_drain(s));
  long long total = 0;
  for (int k = 0; k < sm_stream_device_count(s); k++) {
    printf("worker %d: %lld frames\n", k, sm_stream_frames_done(s, k));
    total += sm_stream_frames_done(s, k);
  }
  printf("frames %lld launches %lld devices %d\n", total, sm_stream_launch_count(s), ndev);
  CHECK(sm_stream_destroy(s));
  f = fopen(argv[2], "wb");
  fwrite(out, 2, npix * n, f);
  fclose(f);
  free(in); free(out); free(tickets);
  return total == n ? 0 : 5;
}
