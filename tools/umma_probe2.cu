// umma_probe2: CTA-pair (cluster of 2, tcgen05 cta_group::2) mechanics against a host GEMM.
//   D[256 x 256] = A[256 x 128] * B[256 x 128]^T ; CTA r holds A rows [128r,128r+128) and B rows (n) [128r,128r+128).
//   Checks: cta_group::2 TMEM alloc in both CTAs, M=256 MMA reading both CTAs' shared memory, multicast commit,
//   peer -> leader remote mbarrier arrive (the "my operands are in place" relay), cluster teardown.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../2024-hl-spi3s-sunerf_b200/csrc/snf_tcgen05.cuh"
using namespace snf::tc;

constexpr int MH = 128, N = 256, NHALF = 128, K = 128;
constexpr int A_IMG = 2 * 128 * 128;      // 2 k-slabs x 128 rows x 128 B
constexpr int B_IMG = 2 * 128 * 128;      // this CTA's half of B

struct Args { const uint8_t *a_img, *b_img; float *d; int use_ts; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2_kernel(Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t rank = cluster_ctarank();
  const uint32_t sA = base, sB = base + A_IMG, sBar = sB + B_IMG;   // bar0: done (multicast), bar1: peer ready (leader only)
  const uint32_t sSlot = sBar + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint8_t *ga = p.a_img + (size_t)rank * A_IMG, *gb = p.b_img + (size_t)rank * B_IMG;
  for (int i = tid; i < A_IMG / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(ga)[i];
  for (int i = tid; i < B_IMG / 16; i += 128) reinterpret_cast<uint4 *>(smem + A_IMG)[i] = reinterpret_cast<const uint4 *>(gb)[i];
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(sBar, 1); mbar_init(sBar + 8, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc_2cta(sSlot, 512);
  tcgen05_fence_before();
  cluster_sync_all();                      // barriers initialised + TMEM allocated in both CTAs
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + A_IMG + B_IMG + 16);
  const uint32_t tm_row = tmem + ((uint32_t)(warp * 32) << 16);
  if (p.use_ts) {   // A into TMEM columns 256.. (each CTA its own 128 rows)
    const int row = warp * 32 + lane;
    for (int kc = 0; kc < K / 32; ++kc) {
      uint32_t v[16];
      for (int c = 0; c < 16; ++c) {
        const int k = kc * 32 + c * 2;
        v[c] = *reinterpret_cast<const uint32_t *>(smem + (k >> 6) * 16384 + sw128_chunk_off(row, (k & 63) >> 3) + (k & 7) * 2);
      }
      tmem_st16(tm_row + 256 + kc * 16, v);
    }
    tmem_st_wait();
    tcgen05_fence_before();
  }
  __syncthreads();
  if (rank == 1 && tid == 0) mbar_arrive_remote(mapa_shared(sBar + 8, 0));   // "peer operands are in place"
  if (rank == 0 && tid == 0) {
    mbar_wait_cluster(sBar + 8, 0);
    tcgen05_fence_after();
    const uint32_t idesc = idesc_bf16(256, N);
    for (int ks = 0; ks < 2; ++ks)
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint64_t bd = smem_desc(sB + ks * 16384 + k4 * 32, 16, 1024);
        if (p.use_ts) mma_ts_2cta(tmem, tmem + 256 + (ks * 4 + k4) * 8, bd, idesc, (ks | k4) != 0);
        else mma_ss_2cta(tmem, smem_desc(sA + ks * 16384 + k4 * 32, 16, 1024), bd, idesc, (ks | k4) != 0);
      }
    mma_commit_2cta(sBar, 3);
  }
  mbar_wait(sBar, 0);
  tcgen05_fence_after();
  const int row = rank * MH + warp * 32 + lane;
  for (int g = 0; g < N / 32; ++g) {
    uint32_t acc[32];
    tmem_ld32(tm_row + g * 32, acc);
    tmem_ld_wait(acc);
    for (int i = 0; i < 32; ++i) p.d[row * N + g * 32 + i] = __uint_as_float(acc[i]);
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc_2cta(tmem, 512); }
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFF + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

int main(int argc, char **argv) {
  const int use_ts = argc > 1 && !strcmp(argv[1], "ts");
  const int M = 2 * MH;
  std::vector<float> A(M * K), B(N * K);
  srand(4321);
  for (auto &v : A) v = bf2f(f2bf((float)(rand() % 2001 - 1000) / 1000.f));
  for (auto &v : B) v = bf2f(f2bf((float)(rand() % 2001 - 1000) / 1000.f));
  std::vector<uint8_t> a_img(2 * A_IMG, 0), b_img(2 * B_IMG, 0);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
    uint16_t v = f2bf(A[m * K + k]);
    memcpy(&a_img[(m / MH) * A_IMG + (k >> 6) * 16384 + sw128_chunk_off(m % MH, (k & 63) >> 3) + (k & 7) * 2], &v, 2);
  }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
    uint16_t v = f2bf(B[n * K + k]);
    memcpy(&b_img[(n / NHALF) * B_IMG + (k >> 6) * 16384 + sw128_chunk_off(n % NHALF, (k & 63) >> 3) + (k & 7) * 2], &v, 2);
  }
  uint8_t *da, *db; float *dd;
  cudaMalloc(&da, 2 * A_IMG); cudaMalloc(&db, 2 * B_IMG); cudaMalloc(&dd, M * N * 4);
  cudaMemcpy(da, a_img.data(), 2 * A_IMG, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img.data(), 2 * B_IMG, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0, M * N * 4);
  const int smem = A_IMG + B_IMG + 64;
  cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Args p{da, db, dd, use_ts};
  probe2_kernel<<<2, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("PROBE2 %s: CUDA error %s\n", use_ts ? "ts" : "ss", cudaGetErrorString(e)); return 1; }
  std::vector<float> D(M * N);
  cudaMemcpy(D.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0, bad_lo = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0;
    for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
    double err = fabs(ref - D[m * N + n]);
    if (err > maxerr) maxerr = err;
    if (err > 1e-3) { ++bad; if (m < MH) ++bad_lo; }
  }
  printf("PROBE2 %s: max abs err %.3e, mismatches %d / %d (rows<128: %d) -> %s\n", use_ts ? "ts" : "ss", maxerr, bad, M * N, bad_lo,
         bad == 0 ? "PASS" : "FAIL");
  return bad ? 1 : 0;
}
