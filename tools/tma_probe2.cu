// tma_probe2.cu — minimal TMA sanity variants (which form works on this box?)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BYTES>
__global__ void k_guide(const __grid_constant__ CUtensorMap tensor_map, int x, int y, unsigned* out)
{
  __shared__ alignas(128) uint8_t smem_buffer[BYTES];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, BYTES);
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  unsigned s = 0;
  for (int i = threadIdx.x; i < BYTES; i += blockDim.x) s += smem_buffer[i];
  atomicAdd(out, s);
}

int main(int argc, char** argv)
{
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 4096, H = 2320;
  uint8_t* d;
  CK(cudaMalloc(&d, (size_t)W * H));
  CK(cudaMemset(d, 1, (size_t)W * H));
  EncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
  printf("variant %d: entry point %p query result %d\n", variant, (void*)enc, (int)qr);
  unsigned* out;
  CK(cudaMalloc(&out, 4));
  CK(cudaMemset(out, 0, 4));
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  CUresult r;
  int x = 0, y = 0, bytes = 0;
  if (variant == 0 || variant == 1 || variant == 2) {  // uint8 box 32x23; coords (0,0) | (16,5) | (3,5)
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {32, 23}, es[2] = {1, 1};
    r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    bytes = 32 * 23;
    if (variant == 1) { x = 16; y = 5; }
    if (variant == 2) { x = 3; y = 5; }
  } else if (variant == 3 || variant == 4) {  // int32 elements, box 16x16 (64 B rows)
    cuuint64_t dims[2] = {(cuuint64_t)W / 4, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {16, 16}, es[2] = {1, 1};
    r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    bytes = 64 * 16;
    if (variant == 4) { x = 5; y = 7; }
  } else {  // uint8 box 64x16 with 128-byte L2 promotion
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {64, 16}, es[2] = {1, 1};
    r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    bytes = 64 * 16;
    x = 7; y = 9;
  }
  printf("encode rc %d; descriptor words:", (int)r);
  for (int i = 0; i < 16; i++) printf(" %016llx", (unsigned long long)((uint64_t*)&m)[i]);
  printf("\n");
  if (bytes == 32 * 23) k_guide<32 * 23><<<4, 128>>>(m, x, y, out);
  else k_guide<64 * 16><<<4, 128>>>(m, x, y, out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned h = 0;
  if (e == cudaSuccess) CK(cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost));
  printf("variant %d: %s, sum %u (expect %u)\n", variant, cudaGetErrorString(e), h, 4u * bytes);
  return 0;
}
