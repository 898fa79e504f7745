// Probe: can the chip be split into two SM partitions (CUDA green contexts) so that the encoder of the NEXT image
// batch runs on a small partition while the persistent decode step of the CURRENT batch owns the rest?
//   * split 148 SMs into <small> + remainder, one green context + stream each (driver entry points through
//     cudaGetDriverEntryPoint: the product library must not link libcuda)
//   * runtime-API launches on the green-context streams: are the CTAs confined to the partition (%smid)?
//   * do kernels of the two partitions run concurrently?
//   * cooperative launch with 221 KB of dynamic shared memory on the large partition: grid = its SM count works?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o greenctx_probe greenctx_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("driver error %d at %s:%d (%s)\n", (int)r_, __FILE__, __LINE__, #x); exit(1); } } while (0)

template <typename F>
F entry(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult st;
  CK(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st));
  if (st != cudaDriverEntryPointSuccess || !fn) { printf("no driver entry point %s\n", name); exit(1); }
  return reinterpret_cast<F>(fn);
}

__global__ void smid_kernel(int* out, long long spin) {
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) out[blockIdx.x] = (int)smid;
  const long long t0 = clock64();
  while (clock64() - t0 < spin) { }
}

__global__ void coop_kernel(int* out, unsigned* ctr) {
  extern __shared__ unsigned char sm[];
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  sm[threadIdx.x] = (unsigned char)threadIdx.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    out[blockIdx.x] = (int)smid;
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    while (*(volatile unsigned*)ctr < gridDim.x) {  // software grid barrier: needs every CTA resident
      if (clock64() - t0 > 2000000000LL) { out[blockIdx.x] = -1; break; }
    }
  }
}

int main(int argc, char** argv) {
  const int small = argc > 1 ? atoi(argv[1]) : 16;
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  auto pGetRes = entry<CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType)>("cuDeviceGetDevResource");
  auto pSplit = entry<CUresult (*)(CUdevResource*, unsigned*, const CUdevResource*, CUdevResource*, unsigned, unsigned)>("cuDevSmResourceSplitByCount");
  auto pDesc = entry<CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned)>("cuDevResourceGenerateDesc");
  auto pCreate = entry<CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned)>("cuGreenCtxCreate");
  auto pStream = entry<CUresult (*)(CUstream*, CUgreenCtx, unsigned, int)>("cuGreenCtxStreamCreate");
  auto pDevGet = entry<CUresult (*)(CUdevice*, int)>("cuDeviceGet");

  CUdevice dev;
  CU(pDevGet(&dev, 0));
  CUdevResource all, grp[1], rest;
  CU(pGetRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs: %u\n", all.sm.smCount);
  unsigned n = 1;
  CU(pSplit(grp, &n, &all, &rest, 0, (unsigned)small));
  printf("split: %u group(s) of %u SMs, remainder %u SMs\n", n, grp[0].sm.smCount, rest.sm.smCount);
  CUdevResourceDesc d_small, d_rest;
  CU(pDesc(&d_small, &grp[0], 1));
  CU(pDesc(&d_rest, &rest, 1));
  CUgreenCtx g_small, g_rest;
  CU(pCreate(&g_small, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CU(pCreate(&g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream s_small, s_rest;
  CU(pStream(&s_small, g_small, CU_STREAM_NON_BLOCKING, 0));
  CU(pStream(&s_rest, g_rest, CU_STREAM_NON_BLOCKING, 0));

  const int NB = 2000;
  int *o_small, *o_rest;
  CK(cudaMalloc(&o_small, NB * sizeof(int)));
  CK(cudaMalloc(&o_rest, NB * sizeof(int)));
  // 1. confinement
  smid_kernel<<<NB, 64, 0, (cudaStream_t)s_small>>>(o_small, 2000);
  smid_kernel<<<NB, 64, 0, (cudaStream_t)s_rest>>>(o_rest, 2000);
  CK(cudaDeviceSynchronize());
  std::vector<int> h(NB);
  std::set<int> a, b;
  CK(cudaMemcpy(h.data(), o_small, NB * sizeof(int), cudaMemcpyDeviceToHost));
  for (int v : h) a.insert(v);
  CK(cudaMemcpy(h.data(), o_rest, NB * sizeof(int), cudaMemcpyDeviceToHost));
  for (int v : h) b.insert(v);
  int common = 0;
  for (int v : a) common += b.count(v);
  printf("runtime launches on green-context streams: small partition used %zu SMs, large %zu SMs, %d in common\n", a.size(), b.size(), common);

  // 2. concurrency: one long CTA per SM of each partition, alone and together
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const long long spin = 4000000;  // ~2 ms
  auto timed = [&](bool do_small, bool do_rest) {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, 0));
    CK(cudaStreamWaitEvent((cudaStream_t)s_small, e0, 0));
    CK(cudaStreamWaitEvent((cudaStream_t)s_rest, e0, 0));
    if (do_small) smid_kernel<<<grp[0].sm.smCount, 64, 0, (cudaStream_t)s_small>>>(o_small, spin);
    if (do_rest) smid_kernel<<<rest.sm.smCount, 64, 0, (cudaStream_t)s_rest>>>(o_rest, spin);
    cudaEvent_t j0, j1;
    CK(cudaEventCreateWithFlags(&j0, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&j1, cudaEventDisableTiming));
    CK(cudaEventRecord(j0, (cudaStream_t)s_small));
    CK(cudaEventRecord(j1, (cudaStream_t)s_rest));
    CK(cudaStreamWaitEvent(0, j0, 0));
    CK(cudaStreamWaitEvent(0, j1, 0));
    CK(cudaEventRecord(e1, 0));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
  };
  timed(true, true);
  printf("spin kernels: small alone %.3f ms, large alone %.3f ms, both %.3f ms\n", timed(true, false), timed(false, true), timed(true, true));

  // 3. cooperative launch with 221 KB of shared memory on the large partition
  const int smem = 221 * 1024;
  CK(cudaFuncSetAttribute(coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  unsigned* ctr;
  CK(cudaMalloc(&ctr, sizeof(unsigned)));
  for (int grid : {(int)rest.sm.smCount, (int)all.sm.smCount}) {
    CK(cudaMemset(ctr, 0, sizeof(unsigned)));
    CK(cudaMemset(o_rest, 0xff, NB * sizeof(int)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)s_rest;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // keep the small partition busy meanwhile
    smid_kernel<<<grp[0].sm.smCount * 4, 64, 0, (cudaStream_t)s_small>>>(o_small, spin);
    cudaError_t e = cudaLaunchKernelEx(&cfg, coop_kernel, o_rest, ctr);
    if (e != cudaSuccess) {
      printf("cooperative launch, grid %d on the large partition: launch error %s\n", grid, cudaGetErrorString(e));
      cudaGetLastError();
      CK(cudaDeviceSynchronize());
      continue;
    }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cooperative launch, grid %d: %s\n", grid, cudaGetErrorString(e)); break; }
    CK(cudaMemcpy(h.data(), o_rest, grid * sizeof(int), cudaMemcpyDeviceToHost));
    std::set<int> c;
    int timeouts = 0, shared = 0;
    for (int i = 0; i < grid; ++i) { if (h[i] < 0) ++timeouts; else c.insert(h[i]); }
    for (int v : c) shared += a.count(v);
    printf("cooperative launch, grid %d, 221 KB smem on the large partition: %zu distinct SMs, %d barrier time-outs, %d SMs of the small partition used\n",
           grid, c.size(), timeouts, shared);
  }
  // 4. same cooperative kernel on the legacy (whole-device) stream while the small partition is busy
  {
    CK(cudaMemset(ctr, 0, sizeof(unsigned)));
    cudaStream_t plain;
    CK(cudaStreamCreateWithFlags(&plain, cudaStreamNonBlocking));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(rest.sm.smCount);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = plain;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    smid_kernel<<<grp[0].sm.smCount * 4, 64, 0, (cudaStream_t)s_small>>>(o_small, spin);
    CK(cudaLaunchKernelEx(&cfg, coop_kernel, o_rest, ctr));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), o_rest, rest.sm.smCount * sizeof(int), cudaMemcpyDeviceToHost));
    int shared = 0;
    for (unsigned i = 0; i < rest.sm.smCount; ++i) shared += a.count(h[i]);
    printf("cooperative launch on a PLAIN stream (grid %u) while the small partition is busy: %d CTAs landed on the small partition's SMs\n", rest.sm.smCount, shared);
  }
  printf("ok\n");
  return 0;
}
