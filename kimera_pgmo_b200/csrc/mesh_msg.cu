// KimeraPgmoMesh wire format <-> the device-resident cloud (SURVEY 8f "next" row 4).
//
// kimera_pgmo_msgs/msg/KimeraPgmoMesh.msg carries geometry_msgs/Point[] vertices (3 doubles),
// time[] vertex_stamps ({uint32 sec, uint32 nsec}), std_msgs/ColorRGBA[] vertex_colors (4 floats)
// and int32[] vertex_indices.  conversions::fillFromMsg / fillMsg
// (kimera_pgmo_ros/include/kimera_pgmo_ros/conversion/mesh_conversion.h:69-199) convert it vertex
// by vertex on the host; here the message arrays go to the GPU as they are and are converted
// there, into the 16-byte {x, y, z, rgba} records + u64 stamps that pgmo_deform_mesh_device takes.
#include "deform_kernels.cuh"

namespace pgmo {

namespace {

// fillFromMsg (:153-199): Pos(p.x, p.y, p.z) narrows double -> float (round to nearest); colour
// byte = static_cast<uint8_t>(255.0f * c) (truncation); stamp = Time::toNSec() = sec * 1e9 + nsec;
// without colours the point keeps PointXYZRGBA's default (0, 0, 0, 255).
__global__ void msgToCloudKernel(uint64_t n, const double* __restrict__ pts,
                                 const float* __restrict__ colors,
                                 const uint32_t* __restrict__ stamps2, float4* __restrict__ cloud,
                                 uint64_t* __restrict__ stamps) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v;
  v.x = __double2float_rn(pts[3 * i + 0]);
  v.y = __double2float_rn(pts[3 * i + 1]);
  v.z = __double2float_rn(pts[3 * i + 2]);
  uint32_t rgba = 0xFF000000u;
  if (colors) {
    const float4 c = reinterpret_cast<const float4*>(colors)[i];
    const uint32_t r = static_cast<uint8_t>(static_cast<int>(__fmul_rn(255.0f, c.x)));
    const uint32_t g = static_cast<uint8_t>(static_cast<int>(__fmul_rn(255.0f, c.y)));
    const uint32_t b = static_cast<uint8_t>(static_cast<int>(__fmul_rn(255.0f, c.z)));
    const uint32_t a = static_cast<uint8_t>(static_cast<int>(__fmul_rn(255.0f, c.w)));
    rgba = (a << 24) | (r << 16) | (g << 8) | b;
  }
  v.w = __uint_as_float(rgba);
  cloud[i] = v;
  if (stamps && stamps2) {
    stamps[i] = static_cast<uint64_t>(stamps2[2 * i]) * 1000000000ull + stamps2[2 * i + 1];
  }
}

// fillMsg (:69-139): p.x = pos.x() widens float -> double; colour = (1.0f / 255) * byte;
// Time::fromNSec: sec = ns / 1e9, nsec = ns % 1e9.
__global__ void cloudToMsgKernel(uint64_t n, const float4* __restrict__ cloud,
                                 const uint64_t* __restrict__ stamps, double* __restrict__ pts,
                                 float* __restrict__ colors, uint32_t* __restrict__ stamps2) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = cloud[i];
  pts[3 * i + 0] = static_cast<double>(v.x);
  pts[3 * i + 1] = static_cast<double>(v.y);
  pts[3 * i + 2] = static_cast<double>(v.z);
  if (colors) {
    const uint32_t rgba = __float_as_uint(v.w);
    const float conv = __fdiv_rn(1.0f, 255.0f);
    float4 c;
    c.x = __fmul_rn(conv, static_cast<float>((rgba >> 16) & 255u));
    c.y = __fmul_rn(conv, static_cast<float>((rgba >> 8) & 255u));
    c.z = __fmul_rn(conv, static_cast<float>(rgba & 255u));
    c.w = __fmul_rn(conv, static_cast<float>(rgba >> 24));
    reinterpret_cast<float4*>(colors)[i] = c;
  }
  if (stamps && stamps2) {
    const uint64_t ns = stamps[i];
    stamps2[2 * i] = static_cast<uint32_t>(ns / 1000000000ull);
    stamps2[2 * i + 1] = static_cast<uint32_t>(ns % 1000000000ull);
  }
}

}  // namespace

cudaError_t launchMsgToCloud(uint64_t n, const double* pts, const float* colors,
                             const uint32_t* stamps2, float4* cloud, uint64_t* stamps,
                             cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  msgToCloudKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(n, pts, colors, stamps2,
                                                                              cloud, stamps);
  return cudaGetLastError();
}

cudaError_t launchCloudToMsg(uint64_t n, const float4* cloud, const uint64_t* stamps, double* pts,
                             float* colors, uint32_t* stamps2, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  cloudToMsgKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(n, cloud, stamps, pts,
                                                                              colors, stamps2);
  return cudaGetLastError();
}

}  // namespace pgmo

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadMeshMsgKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, msgToCloudKernel);
  cudaFuncGetAttributes(&attr, cloudToMsgKernel);
  cudaGetLastError();
}
}  // namespace pgmo
