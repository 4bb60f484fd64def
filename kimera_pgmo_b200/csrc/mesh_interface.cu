#include "mesh_interface.cuh"

namespace pgmo {

namespace {

// one thread per vertex; the block of a vertex by binary search over the (few) block offsets
__global__ void voxbloxDecodeKernel(uint32_t n_blocks, const int32_t* __restrict__ block_idx,
                                    const uint64_t* __restrict__ offsets, uint64_t n,
                                    const uint16_t* __restrict__ xyz16,
                                    const uint8_t* __restrict__ rgb, float edge,
                                    float* __restrict__ xyz, uint32_t* __restrict__ rgba) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t lo = 0, hi = n_blocks;  // last block with offsets[b] <= i (empty blocks skipped)
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= i) lo = mid;
    else hi = mid;
  }
  // 2.0f / std::numeric_limits<uint16_t>::max() in float (voxblox_msg_interface.h:33-34)
  const float conv = __fdiv_rn(2.0f, 65535.0f);
  // (float(u16) * conv + float(index)) * edge, every operation rounded separately (:46-54)
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float q = static_cast<float>(xyz16[a * n + i]);
    const float idx = static_cast<float>(block_idx[3 * lo + a]);
    xyz[3 * i + a] = __fmul_rn(__fadd_rn(__fmul_rn(q, conv), idx), edge);
  }
  rgba[i] = 0xFF000000u | (static_cast<uint32_t>(rgb[i]) << 16) |
            (static_cast<uint32_t>(rgb[n + i]) << 8) | rgb[2 * n + i];
}

}  // namespace

int stageVoxblox(VoxbloxStage* stage, uint32_t n_blocks, const int32_t* block_idx_xyz,
                 const uint64_t* block_offsets, const VoxbloxInput& in, float* d_xyz,
                 uint32_t* d_rgba, cudaStream_t stream) {
  const uint64_t n = n_blocks ? block_offsets[n_blocks] : 0;
  if (n == 0) return PGMO_OK;
  PGMO_CHECK_ARG(in.x && in.y && in.z && in.r && in.g && in.b);
  PGMO_TRY(stage->xyz16.reserve(3 * n));
  PGMO_TRY(stage->rgb.reserve(3 * n));
  PGMO_TRY(stage->idx.reserve(3ull * n_blocks));
  PGMO_TRY(stage->off.reserve(n_blocks + 1ull));
  const uint16_t* c16[3] = {in.x, in.y, in.z};
  const uint8_t* c8[3] = {in.r, in.g, in.b};
  for (int a = 0; a < 3; ++a) {
    PGMO_CUDA(cudaMemcpyAsync(stage->xyz16.ptr + a * n, c16[a], n * sizeof(uint16_t),
                              cudaMemcpyHostToDevice, stream));
    PGMO_CUDA(cudaMemcpyAsync(stage->rgb.ptr + a * n, c8[a], n, cudaMemcpyHostToDevice, stream));
  }
  PGMO_CUDA(cudaMemcpyAsync(stage->idx.ptr, block_idx_xyz, 3ull * n_blocks * sizeof(int32_t),
                            cudaMemcpyHostToDevice, stream));
  PGMO_CUDA(cudaMemcpyAsync(stage->off.ptr, block_offsets, (n_blocks + 1ull) * sizeof(uint64_t),
                            cudaMemcpyHostToDevice, stream));
  voxbloxDecodeKernel<<<divUp(n, 256), 256, 0, stream>>>(n_blocks, stage->idx.ptr, stage->off.ptr, n,
                                                         stage->xyz16.ptr, stage->rgb.ptr,
                                                         in.block_edge_length, d_xyz, d_rgba);
  PGMO_CUDA(cudaGetLastError());
  return PGMO_OK;
}

}  // namespace pgmo

extern "C" int pgmo_voxblox_decode(int device, uint32_t n_blocks, const int32_t* block_idx_xyz,
                                   const uint64_t* block_offsets, const uint16_t* x,
                                   const uint16_t* y, const uint16_t* z, const uint8_t* r,
                                   const uint8_t* g, const uint8_t* b, float block_edge_length,
                                   float* xyz_out, uint32_t* rgba_out) {
  using namespace pgmo;
  PGMO_CHECK_ARG(n_blocks == 0 || (block_idx_xyz && block_offsets));
  PGMO_TRY(ensureDevice(device));
  const uint64_t n = n_blocks ? block_offsets[n_blocks] : 0;
  if (n == 0) return PGMO_OK;
  PGMO_CHECK_ARG(xyz_out && rgba_out);
  VoxbloxStage stage;
  DevBuf<float> d_xyz;
  DevBuf<uint32_t> d_rgba;
  PGMO_TRY(d_xyz.reserve(3 * n));
  PGMO_TRY(d_rgba.reserve(n));
  const VoxbloxInput in{x, y, z, r, g, b, block_edge_length};
  PGMO_TRY(stageVoxblox(&stage, n_blocks, block_idx_xyz, block_offsets, in, d_xyz.ptr, d_rgba.ptr,
                        nullptr));
  PGMO_CUDA(cudaMemcpy(xyz_out, d_xyz.ptr, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
  PGMO_CUDA(cudaMemcpy(rgba_out, d_rgba.ptr, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return PGMO_OK;
}

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadMeshInterfaceKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, voxbloxDecodeKernel);
  cudaGetLastError();
}
}  // namespace pgmo
