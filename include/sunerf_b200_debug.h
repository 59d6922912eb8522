/*
 * sunerf_b200_debug.h - measurement aids exported by libsunerf_b200.so next to the C ABI of sunerf_b200.h.
 * NOT part of the drop-in boundary: process-global, not thread-safe, used by bench.py's roofline pass only.
 */
#ifndef SUNERF_B200_DEBUG_H
#define SUNERF_B200_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif

/* Measurement aid (bench.py): per-kernel CUDA-event timing of snf_mlp_bwd_bf16 on its launch stream.
 * snf_debug_time_backward(1) arms it and clears the sums, snf_debug_backward_ms(out[3]) returns the number of timed calls
 * and the summed milliseconds of {dgrad chain, wgrad, output-layer gradient}.  Off by default. */
int snf_debug_time_backward(int on);
int snf_debug_backward_ms(double *out_ms);

#ifdef __cplusplus
}
#endif
#endif
