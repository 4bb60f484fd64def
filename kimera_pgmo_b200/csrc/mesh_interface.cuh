// MeshInterface inputs on the device.  The compressors take a message as flat float xyz + rgba
// arrays; a voxblox mesh message carries u16 fixed-point coordinates per block instead
// (VoxbloxMsgInterface, kimera_pgmo_ros/src/voxblox_msg_interface.cpp:43-62).  Decoding it on the
// GPU keeps the per-vertex virtual getActiveVertex() calls off the host and shrinks the upload
// from 16 to 9 bytes per vertex.
#pragma once

#include "common.cuh"

namespace pgmo {

struct VoxbloxInput {
  const uint16_t* x;
  const uint16_t* y;
  const uint16_t* z;
  const uint8_t* r;
  const uint8_t* g;
  const uint8_t* b;
  float block_edge_length;
};

struct VoxbloxStage {
  DevBuf<uint16_t> xyz16;  // x | y | z planes
  DevBuf<uint8_t> rgb;     // r | g | b planes
  DevBuf<int32_t> idx;
  DevBuf<uint64_t> off;
};

// Uploads the packed message (pageable or pinned host memory) and decodes it into d_xyz
// (n x 3 float) / d_rgba (n, PCL packing) on `stream`.
int stageVoxblox(VoxbloxStage* stage, uint32_t n_blocks, const int32_t* block_idx_xyz,
                 const uint64_t* block_offsets, const VoxbloxInput& in, float* d_xyz,
                 uint32_t* d_rgba, cudaStream_t stream);

}  // namespace pgmo
