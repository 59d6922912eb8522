// umma_probe: single-CTA checks of the tcgen05 operand layouts the MLP kernels rely on, against a host GEMM.
//   ss   : A,B K-major SWIZZLE_128B in shared memory (forward / dgrad operands)
//   mn   : A,B MN-major SWIZZLE_128B (wgrad operands: the saved activation images read "transposed")
//   mn2  : same with LBO/SBO swapped (the alternative reading of the descriptor fields)
//   ts   : A from TMEM (packed bf16x2 written by tcgen05.st), B K-major in shared memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../2024-hl-spi3s-sunerf_b200/csrc/snf_tcgen05.cuh"

using namespace snf::tc;

constexpr int M = 128, N = 256, K = 128;          // K = 2 slabs of 64
constexpr int A_IMG = 2 * 128 * 128;               // bytes: 2 slabs x 128 rows x 128 B
constexpr int B_IMG = 2 * 256 * 128;

struct ProbeArgs {
  const uint8_t *a_img, *b_img;
  float *d;           // [M][N]
  int mode;           // 0 ss, 1 mn, 2 mn swapped, 3 ts
  int a_fmt, b_fmt;   // 0 = fp16, 1 = bf16 (kind::f16 operand formats, chosen independently)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *g = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + A_IMG, sBar = sB + B_IMG, sSlot = sBar + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < A_IMG / 16; i += 128) reinterpret_cast<uint4 *>(g)[i] = reinterpret_cast<const uint4 *>(p.a_img)[i];
  for (int i = tid; i < B_IMG / 16; i += 128) reinterpret_cast<uint4 *>(g + A_IMG)[i] = reinterpret_cast<const uint4 *>(p.b_img)[i];
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(sBar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(sSlot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(g + A_IMG + B_IMG + 8);
  const uint32_t tm_row = tmem + ((uint32_t)(warp * 32) << 16);

  if (p.mode == 3) {
    // A (row = lane, k packed two per 32-bit column) -> TMEM columns 256..319 ; rows come from the K-major image
    const int row = warp * 32 + lane;
    for (int kc = 0; kc < K / 32; ++kc) {   // 16 columns (32 k) per store
      uint32_t v[16];
      for (int c = 0; c < 16; ++c) {
        const int k = kc * 32 + c * 2;
        const int slab = k >> 6, kk = k & 63;
        v[c] = *reinterpret_cast<const uint32_t *>(g + slab * 16384 + sw128_chunk_off(row, kk >> 3) + (kk & 7) * 2);
      }
      tmem_st16(tm_row + 256 + kc * 16, v);
    }
    tmem_st_wait();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
  }

  if (tid == 0) {
    if (p.mode == 0 || p.mode == 3) {
      const uint32_t idesc = idesc_f16kind(M, N, p.a_fmt, p.b_fmt);
      for (int ks = 0; ks < 2; ++ks)
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint64_t bd = smem_desc(sB + ks * (256 * 128) + k4 * 32, 16, 1024);
          if (p.mode == 0) {
            const uint64_t ad = smem_desc(sA + ks * 16384 + k4 * 32, 16, 1024);
            mma_ss(tmem, ad, bd, idesc, (ks | k4) != 0);
          } else {
            mma_ts(tmem, tmem + 256 + (ks * 4 + k4) * 8, bd, idesc, (ks | k4) != 0);
          }
        }
    } else {
      // MN-major: image = [slab of 64 MN-elements][k line (128 B)] ; one K=16 step = 16 lines = 2 KB
      const uint32_t idesc = idesc_f16kind(M, N, p.a_fmt, p.b_fmt, 1, 1);
      const uint32_t lbo_a = p.mode == 1 ? 16384u : 1024u, sbo_a = p.mode == 1 ? 1024u : 16384u;
      const uint32_t lbo_b = p.mode == 1 ? 16384u : 1024u, sbo_b = p.mode == 1 ? 1024u : 16384u;
      for (int k16 = 0; k16 < K / 16; ++k16) {
        const uint64_t ad = smem_desc(sA + k16 * 2048, lbo_a, sbo_a);
        const uint64_t bd = smem_desc(sB + k16 * 2048, lbo_b, sbo_b);
        mma_ss(tmem, ad, bd, idesc, k16 != 0);
      }
    }
    mma_commit(sBar);
  }
  mbar_wait(sBar, 0);
  tcgen05_fence_after();
  const int row = warp * 32 + lane;
  for (int g32 = 0; g32 < N / 32; ++g32) {
    uint32_t acc[32];
    tmem_ld32(tm_row + g32 * 32, acc);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) p.d[row * N + g32 * 32 + i] = __uint_as_float(acc[i]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

static uint16_t f2bf(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t f2h(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t b) { __half h; memcpy(&h, &b, 2); return __half2float(h); }
static int g_afmt = 1, g_bfmt = 1;
static uint16_t enc(float f, int fmt) { return fmt ? f2bf(f) : f2h(f); }
static float dec(uint16_t b, int fmt) { return fmt ? bf2f(b) : h2f(b); }

int main(int argc, char **argv) {
  const char *name = argc > 1 ? argv[1] : "ss";
  int mode = !strcmp(name, "ss") ? 0 : !strcmp(name, "mn") ? 1 : !strcmp(name, "mn2") ? 2 : !strcmp(name, "ts") ? 3 : -1;
  if (mode < 0) { printf("unknown test %s\n", name); return 2; }
  // optional operand formats: "bf16" (default), "f16", "mixed" (A bf16 x B fp16, the backward kernels' combination)
  const char *fmt = argc > 2 ? argv[2] : "bf16";
  if (!strcmp(fmt, "f16")) { g_afmt = 0; g_bfmt = 0; } else if (!strcmp(fmt, "mixed")) { g_afmt = 1; g_bfmt = 0; }
  else if (!strcmp(fmt, "mixed2")) { g_afmt = 0; g_bfmt = 1; }
  std::vector<float> A(M * K), B(N * K);
  srand(1234);
  for (auto &v : A) v = dec(enc((float)(rand() % 2001 - 1000) / 1000.f, g_afmt), g_afmt);
  for (auto &v : B) v = dec(enc((float)(rand() % 2001 - 1000) / 1000.f, g_bfmt), g_bfmt);
  std::vector<uint8_t> a_img(A_IMG, 0), b_img(B_IMG, 0);
  if (mode == 0 || mode == 3) {
    // K-major: slab ks (64 k), row r, chunk c8 -> sw128 ; B has 256 rows per slab
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) {
      uint16_t v = enc(A[r * K + k], g_afmt);
      memcpy(&a_img[(k >> 6) * 16384 + sw128_chunk_off(r, (k & 63) >> 3) + (k & 7) * 2], &v, 2);
    }
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) {
      uint16_t v = enc(B[r * K + k], g_bfmt);
      memcpy(&b_img[(k >> 6) * (256 * 128) + sw128_chunk_off(r, (k & 63) >> 3) + (k & 7) * 2], &v, 2);
    }
  } else {
    // MN-major: slab = 64 consecutive m (or n); inside a slab line k holds the 64 values, chunk order swizzled by k%8.
    // This is byte-identical to a K-major image of the TRANSPOSED matrix [k rows][m cols] with 128 rows per slab.
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
      uint16_t v = enc(A[m * K + k], g_afmt);
      memcpy(&a_img[(m >> 6) * 16384 + sw128_chunk_off(k, (m & 63) >> 3) + (m & 7) * 2], &v, 2);
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
      uint16_t v = enc(B[n * K + k], g_bfmt);
      memcpy(&b_img[(n >> 6) * 16384 + sw128_chunk_off(k, (n & 63) >> 3) + (n & 7) * 2], &v, 2);
    }
  }
  uint8_t *da, *db; float *dd;
  cudaMalloc(&da, A_IMG); cudaMalloc(&db, B_IMG); cudaMalloc(&dd, M * N * 4);
  cudaMemcpy(da, a_img.data(), A_IMG, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img.data(), B_IMG, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0, M * N * 4);
  const int smem = A_IMG + B_IMG + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ProbeArgs p{da, db, dd, mode, g_afmt, g_bfmt};
  probe_kernel<<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("PROBE %s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<float> D(M * N);
  cudaMemcpy(D.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0;
    for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
    double err = fabs(ref - D[m * N + n]);
    if (err > maxerr) maxerr = err;
    if (err > 1e-3) ++bad;
  }
  printf("PROBE %s/%s: max abs err %.3e, mismatches %d / %d -> %s\n", name, fmt, maxerr, bad, M * N, bad == 0 ? "PASS" : "FAIL");
  if (bad) {
    printf("  D[0][0..7]  :"); for (int i = 0; i < 8; ++i) printf(" %8.4f", D[i]); printf("\n  ref[0][0..7]:");
    for (int n = 0; n < 8; ++n) { double r = 0; for (int k = 0; k < K; ++k) r += (double)A[k] * B[n * K + k]; printf(" %8.4f", r); }
    printf("\n");
  }
  return bad ? 1 : 0;
}
