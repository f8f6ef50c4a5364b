// Random-gather throughput of one B200 through the paths a traversal kernel can use for its node / primitive records:
// LDG.32/64/128/256 (L1 LSU data pipe) and tex1Dfetch<float4> (L1 TEX data pipe), every lane at its own random record
// of an array that fits L2.  Prints lane-loads/s and bytes/s.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a gather.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned rng(unsigned &s) { s = s * 1664525u + 1013904223u; return s; }

template <int MODE>  // 0: LDG.32, 1: LDG.64, 2: LDG.128, 3: LDG.256, 4: tex float4, 5: 2 x LDG.128 (32 B), 6: LDG.128 + tex float4
__global__ void __launch_bounds__(256) k(const char *base, cudaTextureObject_t tex, unsigned nrec, int iters, float *out) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  unsigned idx = rng(s) % nrec;
  for (int i = 0; i < iters; ++i) {
    const char *p = base + (size_t)idx * 128u;
    float v;
    if (MODE == 0) v = __ldg((const float *)p);
    if (MODE == 1) { float2 a = __ldg((const float2 *)p); v = a.x + a.y; }
    if (MODE == 2) { float4 a = __ldg((const float4 *)p); v = a.x + a.y + a.z + a.w; }
    if (MODE == 3) {
      float r[8];
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
      v = r[0] + r[1] + r[2] + r[3] + r[4] + r[5] + r[6] + r[7];
    }
    if (MODE == 4) { float4 a = tex1Dfetch<float4>(tex, (int)(idx * 8u)); v = a.x + a.y + a.z + a.w; }
    if (MODE == 5) { float4 a = __ldg((const float4 *)p), b = __ldg((const float4 *)(p + 16)); v = a.x + a.y + a.z + a.w + b.x + b.w; }
    if (MODE == 6) { float4 a = __ldg((const float4 *)p), b = tex1Dfetch<float4>(tex, (int)(idx * 8u + 1u)); v = a.x + a.y + a.z + a.w + b.x + b.w; }
    acc += v;
    idx = (rng(s) + (unsigned)(int)(v * 1e-30f)) % nrec;  // (the next address depends on the load: one round trip per iteration, like a traversal)
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, int bytes_per_load, const char *base, cudaTextureObject_t tex, unsigned nrec, float *out, int blocks_per_sm) {
  const int iters = 2000, sms = 148;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<MODE><<<sms * blocks_per_sm, 256>>>(base, tex, nrec, 200, out);
  CK(cudaEventRecord(e0));
  k<MODE><<<sms * blocks_per_sm, 256>>>(base, tex, nrec, iters, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  double loads = (double)sms * blocks_per_sm * 256 * iters;
  printf("%-28s %d blocks/SM: %.3f ms  %.1f G lane-loads/s  %.2f per SM per cycle @1.9GHz  %.2f TB/s useful\n", name, blocks_per_sm, ms,
         loads / ms * 1e-6, loads / ms * 1e-6 / 148.0 / 1.9, loads * bytes_per_load / ms * 1e-9);
}

int main(int argc, char **argv) {
  const size_t mb = argc > 1 ? atoi(argv[1]) : 48;  // array size in MB (L2-resident by default)
  const unsigned nrec = (unsigned)(mb * 1024 * 1024 / 128);
  char *base; float *out;
  CK(cudaMalloc(&base, (size_t)nrec * 128));
  CK(cudaMemset(base, 0, (size_t)nrec * 128));
  CK(cudaMalloc(&out, 148 * 8 * 256 * sizeof(float)));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = base;
  rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = (size_t)nrec * 128;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  printf("array %zu MB, %u records of 128 B, one dependent random load per lane per iteration\n", mb, nrec);
  for (int b : {4, 8}) {
    run<0>("LDG.32", 4, base, tex, nrec, out, b);
    run<1>("LDG.64", 8, base, tex, nrec, out, b);
    run<2>("LDG.128", 16, base, tex, nrec, out, b);
    run<3>("LDG.256", 32, base, tex, nrec, out, b);
    run<5>("2 x LDG.128 (32 B)", 32, base, tex, nrec, out, b);
    run<4>("tex1Dfetch float4", 16, base, tex, nrec, out, b);
    run<6>("LDG.128 + tex float4 (32 B)", 32, base, tex, nrec, out, b);
  }
  return 0;
}
