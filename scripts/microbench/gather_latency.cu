// Latency of a dependent, lane-divergent gather from a small array (fits L1) and a large one (L2), one warp per SM:
// does a 256-bit load allocate in L1?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a gather_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
__device__ __forceinline__ unsigned rng(unsigned &s) { s = s * 1664525u + 1013904223u; return s; }
#define LD8(OP, p, r) asm volatile(OP " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p))
template <int MODE>
__global__ void k(const char *base, unsigned nrec, int iters, float *out, long long *cycles) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  unsigned idx = rng(s) % nrec;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const char *p = base + (size_t)idx * 128u;
    float v, r[8];
    if (MODE == 0) { float4 a = __ldg((const float4 *)p); v = a.x + a.w; }
    if (MODE == 1) { LD8("ld.global.nc.v8.f32", p, r); v = r[0] + r[7]; }
    if (MODE == 2) { LD8("ld.global.ca.v8.f32", p, r); v = r[0] + r[7]; }
    if (MODE == 3) { LD8("ld.global.nc.L1::evict_last.v8.f32", p, r); v = r[0] + r[7]; }
    if (MODE == 4) { float4 a; asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p)); v = a.x + a.w; }
    acc += v;
    idx = (rng(s) + (unsigned)(int)(v * 1e-30f)) % nrec;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char *name, const char *base, unsigned nrec, float *out, long long *cyc) {
  const int iters = 4000;
  k<MODE><<<148, 32>>>(base, nrec, 2000, out, cyc);  // warm (fills the caches)
  k<MODE><<<148, 32>>>(base, nrec, iters, out, cyc);
  CK(cudaDeviceSynchronize());
  long long h[148]; CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
  double m = 0; for (int i = 0; i < 148; ++i) m += (double)h[i]; m /= 148.0 * iters;
  printf("  %-36s %.0f cycles per dependent divergent load\n", name, m);
}
int main() {
  char *base; float *out; long long *cyc;
  CK(cudaMalloc(&base, 64u << 20)); CK(cudaMemset(base, 0, 64u << 20));
  CK(cudaMalloc(&out, 148 * 32 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
  for (unsigned kb : {32u, 65536u}) {
    unsigned nrec = kb * 1024u / 128u;
    printf("array %u KB (%u records of 128 B), one warp per SM:\n", kb, nrec);
    run<0>("LDG.128 (ld.global.nc)", base, nrec, out, cyc);
    run<1>("LDG.256 (ld.global.nc.v8)", base, nrec, out, cyc);
    run<2>("LDG.256 (ld.global.ca.v8)", base, nrec, out, cyc);
    run<3>("LDG.256 (nc, L1::evict_last)", base, nrec, out, cyc);
    run<4>("LDG.128 (ld.global.cg: L2 only)", base, nrec, out, cyc);
  }
  return 0;
}
