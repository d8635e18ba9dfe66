// Layout conversion at the API boundary: planar [N][C][S] <-> channels-last [N][S][C].
// 32x32 shared-memory tile transpose: coalesced 128 B reads and writes on both sides (HBM-bound, 8 B/element).
#include "common.cuh"

namespace rcmvs {

// in: [rows][cols] -> out: [cols][rows], batched over blockIdx.z
__global__ void transpose2d_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int64_t cols) {
  __shared__ float tile[32][33];
  const int64_t batch_off = (int64_t)blockIdx.z * rows * cols;
  in += batch_off;
  out += batch_off;
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j;
    int64_t c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[(int64_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t c = c0 + j;
    int r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][j];
  }
}

static int launch_transpose(const float* src, float* dst, int N, int64_t rows, int64_t cols, cudaStream_t st,
                            const char* name) {
  RCMVS_REQUIRE(src && dst, RCMVS_ERR_BAD_ARG, "%s: null pointer", name);
  RCMVS_REQUIRE(N > 0 && rows > 0 && cols > 0, RCMVS_ERR_BAD_ARG, "%s: non-positive dims", name);
  // grid.x (unbounded) walks cols = S, grid.y walks rows = C
  RCMVS_REQUIRE(ceil_div<int64_t>(rows, 32) <= 65535 && N < 65536, RCMVS_ERR_UNSUPPORTED, "%s: dims too large", name);
  dim3 grid((unsigned)ceil_div<int64_t>(cols, 32), (unsigned)ceil_div<int64_t>(rows, 32), N);
  transpose2d_kernel<<<grid, dim3(32, 8), 0, st>>>(src, dst, (int)rows, cols);
  RCMVS_CHECK_LAUNCH(name);
  return RCMVS_OK;
}

}  // namespace rcmvs

extern "C" int rcmvs_planar_to_cl(const float* src, float* dst, int N, int C, int64_t S, rcmvs_stream_t stream) {
  return rcmvs::launch_transpose(src, dst, N, C, S, rcmvs::as_stream(stream), "rcmvs_planar_to_cl");
}

// [N][S][C] -> [N][C][S]: rows = S (long), cols = C (short).  Swap roles so grid.x (unbounded) walks S.
namespace rcmvs {
__global__ void cl_to_planar_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t S, int C) {
  __shared__ float tile[32][33];
  const int64_t batch_off = (int64_t)blockIdx.z * S * C;
  in += batch_off;
  out += batch_off;
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t s = s0 + j;
    int c = c0 + threadIdx.x;
    if (s < S && c < C) tile[j][threadIdx.x] = in[s * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j;
    int64_t s = s0 + threadIdx.x;
    if (s < S && c < C) out[(int64_t)c * S + s] = tile[threadIdx.x][j];
  }
}
}  // namespace rcmvs

extern "C" int rcmvs_cl_to_planar(const float* src, float* dst, int N, int C, int64_t S, rcmvs_stream_t stream) {
  using namespace rcmvs;
  RCMVS_REQUIRE(src && dst, RCMVS_ERR_BAD_ARG, "rcmvs_cl_to_planar: null pointer");
  RCMVS_REQUIRE(N > 0 && C > 0 && S > 0 && N < 65536, RCMVS_ERR_BAD_ARG, "rcmvs_cl_to_planar: bad dims");
  dim3 grid((unsigned)ceil_div<int64_t>(S, 32), (unsigned)ceil_div(C, 32), N);
  cl_to_planar_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src, dst, S, C);
  RCMVS_CHECK_LAUNCH("rcmvs_cl_to_planar");
  return RCMVS_OK;
}
