/* pgmo_b200.h — C ABI of the B200-native Kimera-PGMO mesh path (libpgmo_b200.so).
 *
 * Plain pointers and sizes only; no C++ / torch / PCL / GTSAM types cross this boundary.
 * The header-only C++ shim under include/kimera_pgmo_b200/ re-creates the reference's C++
 * API (same namespaces, names and argument meaning) on top of these entry points; see
 * INTEGRATION.md for the binding a Kimera-PGMO maintainer would add.
 *
 * Reference interfaces replaced (paths relative to the Kimera-PGMO tree):
 *   pgmo_deform_points*      kimera_pgmo/include/kimera_pgmo/mesh_deformation.h:63-75 (unstamped
 *                            deformPoints), :129-141 (stamped deformPoints) incl. interpPoint
 *                            (src/mesh_deformation.cpp:73-106) and SearchTree (:22-70)
 *   pgmo_deform_predeform*   include/kimera_pgmo/deformation_graph.h:779-806 (predeformPoints)
 *   pgmo_vcc_*               include/kimera_pgmo/compression/voxel_clearing_compression.h:14-97,
 *                            src/compression/voxel_clearing_compression.cpp:34-210
 *   pgmo_oct_*               include/kimera_pgmo/compression/octree_compression.h:29-66 +
 *                            mesh_compression.h:85-96 (compressAndIntegrate / pruneStoredMesh of
 *                            the base class, src/compression/mesh_compression.cpp:216-464)
 *   pgmo_stamp_from_sec      src/utils/common_structs.cpp:21-24
 *
 * Conventions
 *   - Return value: 0 = done; > 0 = the reference would have returned early and left the
 *     outputs untouched (so do we); < 0 = error, message via pgmo_last_error().
 *   - Nothing throws across the boundary.  Where the reference would raise
 *     std::out_of_range (.at()), the call returns PGMO_ERR_OUT_OF_RANGE and the shim re-raises.
 *   - Host entry points take host pointers, copy them in, and do not retain them.
 *     *_device entry points take device pointers valid on the handle's device and enqueue on
 *     the given CUDA stream (cudaStream_t passed as void*; NULL = the handle's own stream).
 *     Calls on one handle take effect in call order whatever stream each one was given: a
 *     call that uses another stream than the previous call on the handle first makes its
 *     stream wait (cudaStreamWaitEvent) for what that call queued.  The pose-source buffers
 *     are the exception (their readers are other handles: see below).
 *   - Handles are independent: no global state, one CUDA stream each, never the default
 *     stream; two handles may be driven from two host threads concurrently (the reference
 *     runs two compressors on two std::threads, mesh_frontend_interface.cpp:100-110).  A single
 *     handle is not thread-safe (the reference serialises deformMesh under
 *     KimeraPgmo::interface_mutex_, kimera_pgmo_ros/src/kimera_pgmo.cpp:296-306).
 *   - There is no CPU fallback: every entry point that computes fails with
 *     PGMO_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef PGMO_B200_H_
#define PGMO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGMO_OK 0
#define PGMO_NOOP 1               /* reference early return: outputs untouched */
#define PGMO_ERR_INVALID (-1)     /* bad argument */
#define PGMO_ERR_CUDA (-2)        /* CUDA runtime failure */
#define PGMO_ERR_NO_DEVICE (-3)   /* no usable GPU (there is no CPU fallback) */
#define PGMO_ERR_UNSUPPORTED (-4)
#define PGMO_ERR_OUT_OF_RANGE (-5) /* reference would throw std::out_of_range */

#define PGMO_PRECISION_FP64 0 /* reference arithmetic: f32 distances, f64 weights + blend */
#define PGMO_PRECISION_FP32 1 /* f32 distances (identical associations), f32 weights + blend */

const char* pgmo_last_error(void); /* thread-local, valid until the next call on this thread */
const char* pgmo_version(void);
int pgmo_device_count(void);
uint64_t pgmo_stamp_from_sec(double sec);

/* Page-locked host memory for the caller's staging buffers (what the C++ shim gathers the
 * reference's AoS clouds into): copies from / to it run at PCIe speed and overlap with the
 * kernels; plain pageable memory is accepted everywhere too, only slower.  No reference
 * counterpart (the reference has no device).  bytes == 0 yields *out = NULL. */
int pgmo_host_alloc(uint64_t bytes, void** out);
int pgmo_host_free(void* ptr);

/* ------------------------------------------------------------------------------------------
 * Deformation: deformation::deformPoints behind one handle per (DeformationGraph, prefix).
 * ---------------------------------------------------------------------------------------- */
typedef struct pgmo_deform pgmo_deform_t;

int pgmo_deform_create(int device, pgmo_deform_t** out);
void pgmo_deform_destroy(pgmo_deform_t* h);

/* Control table = what the hot path reads from DeformationGraph::vertex_positions_[prefix],
 * vertex_stamps_[prefix] (deformation_graph.h:678-679) and from `values`:
 *   pos_xyz   n_c x 3 f64 (control_points[j])
 *   stamps_ns n_c u64, may be NULL for the unstamped overload
 *   valid     n_c u8: values.exists(gtsam::Symbol(prefix, j)); NULL = all valid
 *   poses12   n_c x 12 f64: values.at<Pose3>(Symbol(prefix, j)) as row-major R (9) then t (3);
 *             entries with valid[j] == 0 are never read
 * Copies; may be called again after every solve. */
int pgmo_deform_set_controls(pgmo_deform_t* h, uint64_t n_c, const double* pos_xyz,
                             const uint64_t* stamps_ns, const uint8_t* valid,
                             const double* poses12);
/* Same from device memory (e.g. after an NCCL broadcast); stream-ordered on `stream`. */
int pgmo_deform_set_controls_device(pgmo_deform_t* h, uint64_t n_c, const double* d_pos_xyz,
                                    const uint64_t* d_stamps_ns, const uint8_t* d_valid,
                                    const double* d_poses12, void* stream);

/* Replace only the poses (what changes after every solve when no node was added); n_c x 12 f64.
 * The cached window schedule stays valid: it depends on position / stamp / validity only. */
int pgmo_deform_update_poses(pgmo_deform_t* h, const double* poses12);
int pgmo_deform_update_poses_device(pgmo_deform_t* h, const double* d_poses12, void* stream);

/* --- pose source: one upload per solve, every GPU pulls what it needs ----------------------------
 * Multi-GPU deformation replicates the optimised poses.  Instead of a broadcast of the
 * whole table followed by a table-build kernel, one handle OWNS a pose source — two buffers of
 * n_c x 12 doubles in its device's memory, used alternately so that a new solve can be written
 * while slower devices still read the previous one — and every handle (the owner included) builds
 * its derived tables from it with a kernel that reads the source rows directly: a peer device's
 * memory over NVLink, mapped by CUDA IPC (other process) or peer access (same process).  Only the
 * rows [row_begin, row_end) the handle's slab references are pulled
 * (pgmo_deform_slab_control_range).  The caller orders writer and readers (a stream-ordered
 * collective or an event): readers of buffer b must run after the write of buffer b, and the
 * write of buffer b after the readers of its previous content.
 *   create   allocate on the handle's device; ipc_handle64 (optional, 64 bytes) receives the CUDA
 *            IPC handle for other processes, *d_base (optional) the device pointer for handles of
 *            the same process
 *   attach   map the owner's source: pass the IPC handle OR (d_base, owner_device)
 *   write    owner only: poses12 (host or device memory, n_c x 12 f64) -> buffer, stream-ordered
 *   update_poses_from_source   as pgmo_deform_update_poses for rows [row_begin, row_end) */
int pgmo_deform_pose_source_create(pgmo_deform_t* h, uint64_t n_c, void* ipc_handle64,
                                   const double** d_base);
int pgmo_deform_pose_source_attach(pgmo_deform_t* h, uint64_t n_c, const void* ipc_handle64,
                                   const double* d_base, int owner_device);
int pgmo_deform_pose_source_write(pgmo_deform_t* h, int buffer, const double* poses12, void* stream);
int pgmo_deform_update_poses_from_source(pgmo_deform_t* h, int buffer, uint64_t row_begin,
                                         uint64_t row_end, void* stream);
/* The same with the ordering inside the source: sequence numbers (1, 2, ...; buffer = seq & 1)
 * kept in flag words of the source's own memory, written and polled with system-scope release /
 * acquire over NVLink, so that no collective between the processes is needed per solve.
 *   publish   owner: waits on the device until readers 0 .. n_readers - 1 have acknowledged
 *             sequence seq - 2 (the buffer's previous content), copies poses12, marks the buffer
 *             ready for seq.  Stream-ordered, the host does not wait.
 *   update_poses_from_source_seq   reader `reader` (the owner too): a one-thread kernel waits
 *             until the buffer is ready for seq, the table-build kernel pulls rows
 *             [row_begin, row_end), a one-thread kernel acknowledges seq.
 * A wait gives up after 2 s (the pull then reads whatever the buffer holds) and is counted:
 * pgmo_deform_pose_source_timeouts synchronises the handle's stream and returns the count. */
int pgmo_deform_pose_source_publish(pgmo_deform_t* h, uint64_t seq, uint32_t n_readers,
                                    const double* poses12, void* stream);
int pgmo_deform_update_poses_from_source_seq(pgmo_deform_t* h, uint64_t seq, uint32_t reader,
                                             uint64_t row_begin, uint64_t row_end, void* stream);
int pgmo_deform_pose_source_timeouts(pgmo_deform_t* h, uint64_t* n);
/* Control-point rows the slab [first_position, first_position + n_local) can reference under the
 * current schedule: the union of its runs' windows. */
int pgmo_deform_slab_control_range(pgmo_deform_t* h, uint64_t first_position, uint64_t n_local,
                                   uint64_t* row_begin, uint64_t* row_end);

/* Forgets the cached window schedule (and the per-run tables built from it), so that the next
 * pgmo_deform_build_schedule / pgmo_deform_points rebuilds it even if nothing changed.  What a
 * caller gets implicitly whenever control positions / stamps / validity or the mesh stamps
 * change; exposed so that the rebuild can be measured (bench.py "cold" step). */
int pgmo_deform_drop_schedule(pgmo_deform_t* h);

/* Candidate search of the knn kernel: 1 (default) = bucket-ordered candidate lists over each
 * run's alive control points (grid walk for alive sets > 512); 2 = grid walk only (per-run uniform
 * grids, built once per cached window schedule); 0 = brute force over the whole time window.
 * All give bit-identical results (tests/test_deform_gpu.py::test_search_modes_agree). */
int pgmo_deform_set_search_mode(pgmo_deform_t* h, int mode);

/* Optional kernel timing for roofline reporting: when enabled, pgmo_deform_run_device brackets
 * its kernels with CUDA events on the launching stream.  pgmo_deform_last_timings waits for the
 * last run and returns ms2 = [knn+blend kernel, tie resolver kernel] in milliseconds, plus the
 * number of times the window schedule has been (re)built on this handle. */
int pgmo_deform_set_profiling(pgmo_deform_t* h, int enabled);
int pgmo_deform_last_timings(pgmo_deform_t* h, float* ms2, uint64_t* schedule_builds);

/* Association cache.  interpPoint's neighbour set and weights (mesh_deformation.cpp:83-94) depend
 * on the vertex, the control positions / stamps / validity and the window schedule, not on the
 * poses.  The search leaves them in HBM (56 B per vertex at k = 4); a later call whose schedule is
 * unchanged — the usual deformMesh after a new solve — runs the blend kernel alone, which checks a
 * 64-bit fingerprint of every work item's vertex bits against the one the records were computed
 * for and redoes the search for the items that fail it (on the device, stream-ordered).  Results
 * are bit-identical either way.  enabled = 0: every call searches again (default: enabled). */
int pgmo_deform_set_assoc_cache(pgmo_deform_t* h, int enabled);
/* Diagnostics of the last pgmo_deform_run_device (waits for it): ms4 = [search kernel(s), tie
 * resolver, blend kernel(s), all of them] (needs pgmo_deform_set_profiling), counts4 = [work items
 * of the call, items whose records were rejected or missing, vertices that went through the tie
 * resolver, 1 if the call started from cached records]. */
int pgmo_deform_last_step_info(pgmo_deform_t* h, float* ms4, uint64_t* counts4);

/* Diagnostics: the association records of the slab last deformed, 8 u16 per local position in
 * search order: {vertex offset inside its work item, neighbour 0..6 as (control index - run.lo) |
 * bank class << 13}.  *n_slots = how many there are; min(n_slots, capacity) are copied. */
int pgmo_deform_cache_records(pgmo_deform_t* h, uint16_t* records8, uint64_t capacity,
                              uint64_t* n_slots);

/* deformation::deformPoints, host buffers, whole job.
 *   x,y,z        n_v f32 each (SoA vertex stream)
 *   stamps_ns    n_v u64; NULL selects the unstamped overload (mesh_deformation.h:63-113)
 *   subset_idx   optional `indices` argument (n_subset entries, processed in this order)
 *   k, tol_t     as in the reference (k+1 neighbours are blended, the (k+2)-th sets d_max)
 *   out_x/y/z    n_v f32; written only at processed vertices (others keep the caller's data)
 * Optional outputs, one slot per *processing position* p (p < n_subset, or p < n_v):
 *   out_xyz_f64  3 f64: blended point before the final cast<float>() (FP64 mode only)
 *   assoc        assoc_stride u32: neighbour control indices in blend order, 0xFFFFFFFF padded
 *   assoc_cnt    u8: number of blended neighbours (= entries of the reference's std::set,
 *                before set de-duplication)
 *   processed    u8: 1 = vertex deformed (owns one control_point_map entry), 0 = skipped
 *                (mesh_deformation.h:184-186) */
int pgmo_deform_points(pgmo_deform_t* h, uint64_t n_v, const float* x, const float* y,
                       const float* z, const uint64_t* stamps_ns, const uint64_t* subset_idx,
                       uint64_t n_subset, uint32_t k, double tol_t, int precision, float* out_x,
                       float* out_y, float* out_z, double* out_xyz_f64, uint32_t* assoc,
                       uint32_t assoc_stride, uint8_t* assoc_cnt, uint8_t* processed);

/* --- device-resident / sharded path (what bench.py and a multi-GPU caller drive) ---------
 * 1. every rank: pgmo_deform_stamp_stretches_device over its slab of the stamp stream
 * 2. ranks exchange the stretch lists (a few KB) and concatenate them in rank order
 * 3. every rank: pgmo_deform_build_schedule over the global list (identical on all ranks)
 * 4. every rank: pgmo_deform_run_device on its slab [first_position, first_position + n_local)
 */

/* Maximal stretches of equal stamps, in order.  Writes up to `capacity` (stamp, length) pairs
 * to host arrays and the total count to *n_stretches (may exceed capacity: call again). */
int pgmo_deform_stamp_stretches_device(pgmo_deform_t* h, uint64_t n_local,
                                       const uint64_t* d_stamps_ns, uint64_t capacity,
                                       uint64_t* stretch_stamp, uint64_t* stretch_len,
                                       uint64_t* n_stretches, void* stream);

/* K1 without the host round trip, for callers that keep a schedule across solves: _begin queues
 * the stretch scan of the slab's stamps and the read-back of its 64-bit fingerprint on `stream`;
 * _finish waits for that read-back only — work queued behind it (pgmo_deform_run_device on the
 * cached schedule) keeps running — and sets *changed = 0 iff the slab still has the stretch list
 * last returned by pgmo_deform_stamp_stretches_device.  On *changed = 1 the caller discards the
 * speculative result: stamp_stretches_device -> (exchange) -> build_schedule -> run_device.
 * The scan runs on a side stream of the handle forked off `stream` at the call (it sees what was
 * queued on `stream` before, and overlaps with what is queued after): d_stamps_ns must not be
 * written between _begin and the return of _finish. */
int pgmo_deform_stamp_check_begin(pgmo_deform_t* h, uint64_t n_local, const uint64_t* d_stamps_ns,
                                  void* stream);
int pgmo_deform_stamp_check_finish(pgmo_deform_t* h, int* changed);
/* enabled = 0: the scan is queued on `stream` itself, ahead of what follows (its result is then
 * on the host a few tens of microseconds after _begin, whatever runs behind it).  What a caller
 * wants whose host has work to do between _finish and the next step — e.g. the all-reduce of the
 * flag between the ranks of a sharded job: with the overlapped scan (default, a side stream)
 * _finish can return late in the step, and that work then delays the next step's launches
 * (8 ranks: 0.75 ms per step against 0.54 ms).  PGMO_SCAN_PRIORITY=1 (environment, read at
 * pgmo_deform_create) gives the side stream the highest priority instead: the result arrives
 * early again and the step keeps most of the overlap (0.480 ms against 0.470 / 0.496 ms). */
int pgmo_deform_set_scan_overlap(pgmo_deform_t* h, int enabled);

/* Replays the reference's add/remove window state machine (mesh_deformation.h:159-212) over the
 * stamp stretches and uploads the resulting schedule.  stretch_stamp == NULL selects the
 * unstamped overload for `total_positions` vertices.  Returns PGMO_NOOP for the reference's
 * early returns (:78-80, :83-86, :96-99, :143-152). */
int pgmo_deform_build_schedule(pgmo_deform_t* h, uint64_t n_stretches,
                               const uint64_t* stretch_stamp, const uint64_t* stretch_len,
                               uint64_t total_positions, uint32_t k, double tol_t);

/* Deforms positions [first_position, first_position + n_local).  All vertex arrays are local
 * to the slab (element 0 = position first_position); d_subset_idx (optional) holds local
 * vertex indices.  Optional outputs as in pgmo_deform_points (d_assoc/d_assoc_cnt/d_processed/
 * d_out_xyz_f64 may be NULL).  *n_resolved (optional, host) receives the number of vertices
 * that went through the exact tie-order resolver; requesting it synchronises the stream. */
int pgmo_deform_run_device(pgmo_deform_t* h, uint64_t first_position, uint64_t n_local,
                           const float* d_x, const float* d_y, const float* d_z,
                           const uint64_t* d_subset_idx, int precision, float* d_out_x,
                           float* d_out_y, float* d_out_z, double* d_out_xyz_f64,
                           uint32_t* d_assoc, uint32_t assoc_stride, uint8_t* d_assoc_cnt,
                           uint8_t* d_processed, uint64_t* n_resolved, void* stream);

/* stats4 = [schedule builds so far, stretch the last build resumed from (0 = full replay: control
 * points were not merely appended, or nothing could be kept), stretches, runs].  When control
 * points are only appended between two calls (new keyframes) the window schedule is extended
 * from the first stretch that can see them instead of being replayed from the start. */
int pgmo_deform_schedule_stats(pgmo_deform_t* h, uint64_t* stats4);

/* Schedule introspection (tests / diagnostics): number of runs, and per run
 * [first_position, lo, hi, k, skip] as u64 x 5. */
int pgmo_deform_schedule_info(pgmo_deform_t* h, uint64_t* n_runs, uint64_t capacity,
                              uint64_t* runs5);

/* Host buffers in, host buffers out for the slab [first_position, first_position + n_local) on the
 * CURRENT window schedule (pgmo_deform_build_schedule, or the one the last pgmo_deform_points
 * built): what a caller uses when the mesh stamps did not change since the schedule was built —
 * the usual re-deformation after a new solve — and what every rank of a sharded job calls on its
 * slab.  Only the coordinates cross PCIe (12 B per vertex each way instead of 20 + 12); uploads,
 * kernels and downloads are pipelined in slabs of work items.  x / y / z / out_* are local to the
 * slab.  Skipped vertices (mesh_deformation.h:184-186) come back with their input coordinates.
 * Page-locked buffers (pgmo_host_alloc) copy fastest.  Every DMA operation starts on a 4 KB
 * boundary of the arrays (slabs end on work-item boundaries, the copies do not).  Environment
 * knobs, read at pgmo_deform_create: PGMO_PIPE (slabs per call, default 8), PGMO_PIPE_TAPER
 * (shorter first / last slabs), PGMO_COPY_2D (1: when y - x == z - y, i.e. the three arrays are
 * equally spaced planes of one allocation, a chunk of all three moves in one 3-row operation;
 * measured slightly slower than three plain copies, default 0), PGMO_HOST_DIRECT (1: the kernels
 * read and write page-locked buffers themselves, 2: write only; measured slower than the staged
 * copies on B200 / PCIe 5, default 0), PGMO_PIPE_TRACE (prints the timeline of a call). */
int pgmo_deform_run_host_slab(pgmo_deform_t* h, uint64_t first_position, uint64_t n_local,
                              const float* x, const float* y, const float* z, int precision,
                              float* out_x, float* out_y, float* out_z, uint8_t* processed);

/* DeformationGraph::deformMesh (deformation_graph.cpp:523-559) on a mesh that already lives in
 * HBM as the compressors keep it: n_v vertices as 16-byte records {float x, y, z; uint32 rgba}
 * (pgmo_vcc_device_view / pgmo_oct_device_view) plus their u64 stamps (NULL: unstamped
 * overload).  Replaces the blob -> cloud -> blob conversions and the host copies of a-1.  Writes
 * d_out_xyzw[i] = {deformed xyz, input rgba} for every processed vertex and a copy of the input
 * record otherwise (in place allowed); d_processed (optional, n_v bytes) as in
 * pgmo_deform_points.  Stream-ordered on `stream` (NULL: the handle's) apart from the stamp
 * fingerprint read-back; the caller synchronises before reading.  16-byte aligned pointers. */
int pgmo_deform_mesh_device(pgmo_deform_t* h, uint64_t n_v, const void* d_vertices_xyzw,
                            const uint64_t* d_stamps_ns, uint32_t k, double tol_t, int precision,
                            void* d_out_xyzw, uint8_t* d_processed, uint64_t* n_resolved,
                            void* stream);

/* --- every visible GPU behind one call ----------------------------------------------------------
 * A group is DeformationGraph's mesh path over several devices of one box from one host thread's
 * point of view (inside a call the library runs one host thread per device): contiguous vertex
 * slabs, control positions / stamps / validity replicated from the host, the poses written once
 * into device 0's pose source and pulled over NVLink by the other devices' table-build kernels
 * (devices that cannot map device 0 get them from the host), stamp-stretch lists merged so that
 * every device replays the same window schedule.  Results are bit-identical to one device.
 *   create   devices = NULL, n_devices = 0: every visible device
 *   points   as pgmo_deform_points without indices / association outputs
 *   repeat   the same mesh stamps as the last `points` call (a new solve: set_controls or
 *            update_poses in between): only coordinates cross PCIe */
typedef struct pgmo_deform_group pgmo_deform_group_t;
int pgmo_deform_group_create(int n_devices, const int* devices, pgmo_deform_group_t** out);
void pgmo_deform_group_destroy(pgmo_deform_group_t* g);
int pgmo_deform_group_size(pgmo_deform_group_t* g);
int pgmo_deform_group_set_controls(pgmo_deform_group_t* g, uint64_t n_c, const double* pos_xyz,
                                   const uint64_t* stamps_ns, const uint8_t* valid,
                                   const double* poses12);
int pgmo_deform_group_update_poses(pgmo_deform_group_t* g, const double* poses12);
int pgmo_deform_group_points(pgmo_deform_group_t* g, uint64_t n_v, const float* x, const float* y,
                             const float* z, const uint64_t* stamps_ns, uint32_t k, double tol_t,
                             int precision, float* out_x, float* out_y, float* out_z,
                             uint8_t* processed);
int pgmo_deform_group_repeat(pgmo_deform_group_t* g, uint64_t n_v, const float* x, const float* y,
                             const float* z, int precision, float* out_x, float* out_y,
                             float* out_z, uint8_t* processed);

/* KimeraPgmoMesh wire format (kimera_pgmo_msgs/msg/KimeraPgmoMesh.msg) <-> device cloud, i.e.
 * conversions::fillFromMsg / fillMsg (kimera_pgmo_ros/.../conversion/mesh_conversion.h:69-199) on
 * the GPU.  Host arrays are the message's own vectors: vertices_xyz = geometry_msgs/Point[]
 * (3 doubles per vertex), colors_rgba = std_msgs/ColorRGBA[] (4 floats, optional),
 * stamps_sec_nsec = time[] ({uint32 sec, uint32 nsec}, optional).  Device side: the 16-byte
 * {x, y, z, rgba} records + u64 ns stamps of pgmo_deform_mesh_device.
 *   to_cloud: x = float(p.x); byte = uint8(255.0f * c); stamp = sec * 1e9 + nsec; no colours:
 *             (0, 0, 0, 255).  Stream-ordered; the host arrays are read asynchronously (pin them
 *             or synchronise the stream before reusing them).
 *   to_msg:   p.x = double(x); c = (1.0f / 255) * byte; sec = ns / 1e9, nsec = ns % 1e9.
 *             Synchronises the stream before returning.
 * pgmo_deform_mesh_msg = to_cloud -> pgmo_deform_mesh_device -> to_msg: the backend's
 * optimizeFullMesh path from the received message to the message it publishes, vertices only
 * (faces, colours, stamps and vertex_indices of the output message are those of the input). */
int pgmo_deform_msg_to_cloud(pgmo_deform_t* h, uint64_t n_v, const double* vertices_xyz,
                             const float* colors_rgba, const uint32_t* stamps_sec_nsec,
                             void* d_cloud_xyzw, uint64_t* d_stamps_ns, void* stream);
int pgmo_deform_cloud_to_msg(pgmo_deform_t* h, uint64_t n_v, const void* d_cloud_xyzw,
                             const uint64_t* d_stamps_ns, double* vertices_xyz, float* colors_rgba,
                             uint32_t* stamps_sec_nsec, void* stream);
int pgmo_deform_mesh_msg(pgmo_deform_t* h, uint64_t n_v, const double* vertices_xyz,
                         const uint32_t* stamps_sec_nsec, uint32_t k, double tol_t, int precision,
                         double* out_vertices_xyz, uint8_t* processed);

/* DeformationGraph::predeformPoints (deformation_graph.h:779-806): for i >= start_idx with
 * graph_indices[i] >= 0 and a valid pose: out[i] = float(R * (double(v_i) - g) + t); every other
 * i >= start_idx is appended to to_deform (ascending).  Host buffers.
 * Returns PGMO_ERR_OUT_OF_RANGE if a valid pose index is >= n_c (vertex_positions_.at()). */
int pgmo_deform_predeform(pgmo_deform_t* h, uint64_t n_v, const float* x, const float* y,
                          const float* z, const int32_t* graph_indices, uint64_t start_idx,
                          float* out_x, float* out_y, float* out_z, uint64_t* to_deform,
                          uint64_t* n_to_deform);

/* ------------------------------------------------------------------------------------------
 * VoxelClearingCompression: one handle per compressor object.
 * ---------------------------------------------------------------------------------------- */
typedef struct pgmo_vcc pgmo_vcc_t;

int pgmo_vcc_create(double resolution, int device, pgmo_vcc_t** out);
void pgmo_vcc_destroy(pgmo_vcc_t* h);

/* compressAndIntegrate (voxel_clearing_compression.cpp:34-47).  The MeshInterface is passed
 * flattened in blockIndices() order: block b = vertices [block_offsets[b], block_offsets[b+1]),
 * 3 consecutive vertices = 1 face.
 *   block_idx_xyz  n_blocks x 3 i32
 *   xyz            N x 3 f32 (x,y,z interleaved), rgba N u32 (pcl::PointXYZRGBA::rgba)
 *   remap_out      N u64: remapping[block][local] for every input vertex (always overwritten,
 *                  :154)
 *   keys_out       optional N x 3 i64: the voxel keys round(p / resolution) (:134-136) */
int pgmo_vcc_integrate(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                       const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                       double stamp_s, uint64_t* remap_out, int64_t* keys_out);
/* pruneStoredMesh (:57-66) and clearArchivedBlocks (:68-70) */
int pgmo_vcc_prune(pgmo_vcc_t* h, double earliest_s);
int pgmo_vcc_clear_blocks(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz);
/* sizes4 = [num_vertices, num_polygons, num_active, num_invalid] */
int pgmo_vcc_sizes(pgmo_vcc_t* h, uint64_t* sizes4);
/* getVertices / getTimestamps / getStoredPolygons / getActiveVerticesIndex / getInvalidIndices;
 * any pointer may be NULL.  polygons: num_polygons x 3 u32.  invalid is returned ascending
 * (the reference's order follows unordered_set iteration and is not pinned). */
int pgmo_vcc_snapshot(pgmo_vcc_t* h, float* vertices_xyz, uint32_t* vertices_rgba,
                      uint64_t* stamps_ns, uint32_t* polygons, uint64_t* active,
                      uint64_t* invalid);

/* getVertices / getTimestamps without the copy (mesh_compression.h:37-60 copy into caller
 * shared_ptrs): device pointers to all_vertices_ as 16-byte records {float x, y, z; uint32 rgba}
 * and to all_vertex_stamps_ (u64 ns), valid until the next integrate / prune on this handle.
 * Feeds pgmo_deform_mesh_device so that the compressed mesh never leaves HBM. */
int pgmo_vcc_device_view(pgmo_vcc_t* h, const void** d_vertices_xyzw, const uint64_t** d_stamps_ns,
                         uint64_t* n_vertices);

/* Device-resident integrate for throughput measurement: inputs already in HBM, remap written to
 * HBM, stream-ordered, no host synchronisation.  Single-block or multi-block as above;
 * block_offsets / block_idx_xyz stay host arrays (they are tiny). */
int pgmo_vcc_integrate_device(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                              const uint64_t* block_offsets, const float* d_xyz,
                              const uint32_t* d_rgba, double stamp_s, uint64_t* d_remap_out,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * OctreeCompression (MeshCompression base flow + PCL-octree uniqueness tests).
 * ---------------------------------------------------------------------------------------- */
typedef struct pgmo_oct pgmo_oct_t;

int pgmo_oct_create(double resolution, int device, pgmo_oct_t** out);
/* Same MeshCompression base flow with BlockCompression's structure tests
 * (include/kimera_pgmo/compression/block_compression.h:14-58, src/compression/
 * block_compression.cpp:19-81): a vertex is a re-observation iff its cell
 * floor(p / resolution) already holds an active vertex (the first one inserted wins).  Every
 * other pgmo_oct_* entry point works on the handle unchanged. */
int pgmo_oct_create_block(double resolution, int device, pgmo_oct_t** out);
void pgmo_oct_destroy(pgmo_oct_t* h);

/* MeshCompression::compressAndIntegrate (mesh_compression.cpp:216-417).  Inputs as for
 * pgmo_vcc_integrate.  remap_index/remap_kind: dense record of what the reference does to the
 * caller's persistent `remapping` for each input vertex: kind 0 = no write, 1 = overwrite
 * (remapping.at(block)[i] = index, :288), 2 = insert-if-absent (:366-372).  The lists
 * new_indices / new_vertices / new_triangles are fetched with pgmo_oct_last_outputs. */
int pgmo_oct_integrate(pgmo_oct_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                       const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                       double stamp_s, int64_t* remap_index, uint8_t* remap_kind);
/* Failure semantics of the compressors: argument errors are reported before anything changes.  A
 * capacity / range error detected on the device (octree deeper than supported, voxel key outside
 * the packed range, table full) is reported after the stored mesh has started to change; the
 * handle then refuses every further call (PGMO_ERR_INVALID) instead of continuing on a
 * half-updated mesh — create a new compressor. */
int pgmo_oct_prune(pgmo_oct_t* h, double earliest_s);
/* as pgmo_vcc_device_view, for MeshCompression's all_vertices_ / all_vertex_stamps_ */
int pgmo_oct_device_view(pgmo_oct_t* h, const void** d_vertices_xyzw, const uint64_t** d_stamps_ns,
                         uint64_t* n_vertices);
/* sizes8 = [num_vertices, num_polygons, num_active, 0, num_new_vertices, num_new_triangles,
 *           num_new_indices, num_active] */
int pgmo_oct_sizes(pgmo_oct_t* h, uint64_t* sizes8);
int pgmo_oct_snapshot(pgmo_oct_t* h, float* vertices_xyz, uint32_t* vertices_rgba,
                      uint64_t* stamps_ns, uint32_t* polygons, uint64_t* active_index,
                      float* active_xyz, double* active_stamps_s);
int pgmo_oct_last_outputs(pgmo_oct_t* h, float* new_vertices_xyz, uint32_t* new_vertices_rgba,
                          uint32_t* new_triangles, uint64_t* new_indices);
/* Diagnostics of the last pgmo_oct_integrate: stats4 = [fixed-point rounds of the temp-structure
 * pass, vertices that failed the persistent-structure test (candidates), temp vertices created,
 * vertices accepted into the mesh]. */
int pgmo_oct_last_stats(pgmo_oct_t* h, uint64_t* stats4);

/* ------------------------------------------------------------------------------------------
 * DeltaCompression (kimera_pgmo/include/kimera_pgmo/compression/delta_compression.h:60-149,
 * src/compression/delta_compression.cpp): the compressor Hydra drives.  One handle per object.
 * Iteration order of the reference's unordered containers (vertices_map_, block_info_map_,
 * BlockInfo::vertices) is implementation-defined there; here — as in oracle/delta_oracle.h — it
 * is insertion order, so the order of the vertices inside a delta is "first observed first".
 * ---------------------------------------------------------------------------------------- */
typedef struct pgmo_delta pgmo_delta_t;

int pgmo_delta_create(double resolution, int device, pgmo_delta_t** out);
void pgmo_delta_destroy(pgmo_delta_t* h);
/* MeshDelta::Ptr update(MeshInterface&, uint64_t timestamp_ns, HashedIndexMapping* remapping)
 * (:211-259).  Mesh flattened as for pgmo_vcc_integrate (every block a multiple of 3 vertices).
 * remap_out (optional, one per input vertex): the delta index remapping[block][i] receives.
 * The resulting MeshDelta is read with pgmo_delta_sizes / pgmo_delta_fetch. */
int pgmo_delta_update(pgmo_delta_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                      const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                      uint64_t timestamp_ns, uint64_t* remap_out);
/* The same with semantic labels: labels[i] = *mesh.getActiveSemantics(i) of input vertex i
 * (0xFFFFFFFF = std::nullopt).  A voxel keeps the label of the vertex that set its point
 * (delta_compression.cpp:72-77) and hands it to the delta (addVertex, :120, :128, :347).
 * labels == NULL: every vertex of this message is unlabelled. */
int pgmo_delta_update_labeled(pgmo_delta_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                              const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                              const uint32_t* labels, uint64_t timestamp_ns, uint64_t* remap_out);
/* MeshDelta::semantic_updates of the last delta, one entry per vertex_updates entry (0xFFFFFFFF
 * where the reference pushed nothing); *any_label = 1 iff some update of this handle carried
 * labels.  Either pointer may be NULL. */
int pgmo_delta_fetch_labels(pgmo_delta_t* h, uint32_t* labels, int* any_label);
/* archiveBlocksByTime (:306-310) / clearArchivedBlocks (:312-316) */
int pgmo_delta_archive_by_time(pgmo_delta_t* h, uint64_t earliest_time_ns);
int pgmo_delta_archive_blocks(pgmo_delta_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz);
/* sizes12 = [vertex_start, face_start, vertex_updates.size(), archived vertices in the delta,
 * face_updates.size(), face_archive_updates.size(), capacity for prev_to_curr, for deleted, for
 * observed, for new, active voxels, archived_faces_.size()] of the last update. */
int pgmo_delta_sizes(pgmo_delta_t* h, uint64_t* sizes12);
/* The last MeshDelta (mesh_delta.h:65-150): vertex_updates (xyz, rgba) / stamp_updates,
 * face_updates, face_archive_updates (3 u32 each), prev_to_curr as (key, value) pairs in key
 * order, deleted / observed / new indices ascending; the n_* receive the actual counts.  Any
 * pointer may be NULL. */
int pgmo_delta_fetch(pgmo_delta_t* h, float* vertices_xyz, uint32_t* vertices_rgba, uint64_t* stamps_ns,
                     uint32_t* faces, uint32_t* faces_archive, uint64_t* prev_to_curr2,
                     uint64_t* n_prev_to_curr, uint64_t* deleted, uint64_t* n_deleted,
                     uint64_t* observed, uint64_t* n_observed, uint64_t* new_indices, uint64_t* n_new);
/* block_info_map_ in iteration (insertion) order: index, update_time_ns, sequence_number of up to
 * `capacity` blocks; *n_blocks = how many there are.  Lets a caller evaluate an arbitrary
 * archiveBlocks(BlockFilter) (:318-360) and pass the chosen blocks to pgmo_delta_archive_blocks. */
int pgmo_delta_blocks(pgmo_delta_t* h, uint64_t capacity, int32_t* block_idx_xyz,
                      uint64_t* update_time_ns, uint32_t* sequence_number, uint64_t* n_blocks);
/* stats4 = [redundancy fixed-point rounds of the last call, voxel records, blocks, record
 * compactions so far] */
int pgmo_delta_last_stats(pgmo_delta_t* h, uint64_t* stats4);

/* ------------------------------------------------------------------------------------------
 * Mesh -> graph index bookkeeping of MeshFrontendInterface (members msg_to_mesh_idx_,
 * msg_to_graph_idx_, mesh_to_graph_idx_: kimera_pgmo/include/kimera_pgmo/
 * mesh_frontend_interface.h:118-120; updateMeshToGraphMappings:
 * src/mesh_frontend_interface.cpp:197-211).  One handle per MeshFrontendInterface.  The three
 * maps live in HBM: per block a dense msg-index -> stored-index array for each compressor
 * (never cleared, like the reference's per-block maps) and a dense mesh-index -> graph-index
 * array.  Where the reference's result depends on unordered_map iteration order (several msg
 * vertices of one block on the same mesh vertex) the entries are taken in ascending msg index.
 * ---------------------------------------------------------------------------------------- */
typedef struct pgmo_m2g pgmo_m2g_t;

int pgmo_m2g_create(int device, pgmo_m2g_t** out);
void pgmo_m2g_destroy(pgmo_m2g_t* h);
/* One MeshFrontendInterface::update (:100-115) after both compressors have run on the message:
 * applies their writes to the two persistent remappings, then updateMeshToGraphMappings over
 * the message's blocks in the given order.  Mesh flattened as for pgmo_vcc_integrate.
 *   mesh_index / mesh_kind    per input vertex, what the full-mesh compressor did to
 *                             remapping[block][i]: kind 0 = no write, 1 = overwrite, 2 =
 *                             insert-if-absent (pgmo_oct_integrate's remap_index / remap_kind);
 *                             mesh_kind NULL = all overwrite (pgmo_vcc_integrate's remap_out)
 *   graph_index / graph_kind  the same from the graph compressor */
/* Block indices must be unique within a message (PGMO_ERR_UNSUPPORTED otherwise; the same holds
 * for pgmo_vcc_integrate, pgmo_oct_integrate and pgmo_delta_update). */
int pgmo_m2g_update(pgmo_m2g_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                    const uint64_t* block_offsets, const int64_t* mesh_index,
                    const uint8_t* mesh_kind, const int64_t* graph_index,
                    const uint8_t* graph_kind);
/* mesh_to_graph_idx_: its size, and its pairs in ascending mesh index */
int pgmo_m2g_size(pgmo_m2g_t* h, uint64_t* n_pairs);
int pgmo_m2g_fetch(pgmo_m2g_t* h, uint64_t* mesh_idx, uint64_t* graph_idx);
/* msg_to_mesh_idx_[block] (which 0) / msg_to_graph_idx_[block] (which 1,
 * getVoxbloxMsgToGraphMapping, mesh_frontend_interface.h:56) as a dense array over the msg
 * vertex index, -1 = no entry.  *length = its length; min(length, cap) entries are copied.
 * PGMO_ERR_INVALID for a block no update has named (the reference's .at() throws). */
int pgmo_m2g_block_mapping(pgmo_m2g_t* h, const int32_t* block_idx_xyz, int which, int64_t* out,
                           uint64_t cap, uint64_t* length);
/* mesh_to_graph_idx_ in HBM: dense over mesh indices [0, span), -1 = no entry */
int pgmo_m2g_device_view(pgmo_m2g_t* h, const int64_t** d_mesh_to_graph, uint64_t* span);

/* ------------------------------------------------------------------------------------------
 * MeshInterface inputs.  The integrate calls above take what MeshInterface::getActiveVertex
 * yields, flattened.  A voxblox mesh message (VoxbloxMsgInterface,
 * kimera_pgmo_ros/src/voxblox_msg_interface.cpp:11-62) stores u16 fixed-point coordinates per
 * block; the *_voxblox calls take it as it is — per vertex x, y, z (u16) and r, g, b (u8),
 * blocks concatenated in blockIndices() order — and evaluate getActiveVertex on the GPU:
 *   p = (float(u16) * (2.0f / 65535) + float(block index)) * block_edge_length   (float, no FMA)
 *   colour = (r, g, b, 255).
 * ---------------------------------------------------------------------------------------- */
int pgmo_voxblox_decode(int device, uint32_t n_blocks, const int32_t* block_idx_xyz,
                        const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                        const uint16_t* z, const uint8_t* r, const uint8_t* g, const uint8_t* b,
                        float block_edge_length, float* xyz_out, uint32_t* rgba_out);
int pgmo_vcc_integrate_voxblox(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                               const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                               const uint16_t* z, const uint8_t* r, const uint8_t* g,
                               const uint8_t* b, float block_edge_length, double stamp_s,
                               uint64_t* remap_out, int64_t* keys_out);
int pgmo_oct_integrate_voxblox(pgmo_oct_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                               const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                               const uint16_t* z, const uint8_t* r, const uint8_t* g,
                               const uint8_t* b, float block_edge_length, double stamp_s,
                               int64_t* remap_index, uint8_t* remap_kind);

#ifdef __cplusplus
}
#endif
#endif /* PGMO_B200_H_ */
