/* sift3d_b200.h — C ABI of the B200-native 3DSIFT hot path (libsift3d_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, int status returns (0 = ok, non-zero =
 * error; s3d_last_error() gives the message), no exceptions, caller-allocated outputs, opaque
 * handles.  The C++ façade (include/3dsift/cSIFT3D.h, cMatcher.h) and the Python host mirror
 * (3dsift_b200/api.py) are thin callers of these entry points.  There is no CPU fallback: every
 * compute entry point fails with S3D_ERR_CUDA when no sm_100 device is usable.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference/3DSIFT/).
 */
#ifndef SIFT3D_B200_H
#define SIFT3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define S3D_API __declspec(dllexport)
#else
#define S3D_API __attribute__((visibility("default")))
#endif

#define S3D_OK 0
#define S3D_ERR_ARG 1       /* bad argument */
#define S3D_ERR_CUDA 2      /* CUDA runtime / no device */
#define S3D_ERR_CAPACITY 3  /* an internal candidate buffer overflowed */
#define S3D_ERR_STATE 4     /* call order (e.g. get before run) */

#define S3D_DESC_LEN 768    /* Include/cMatcher.h:8 DESC_LENGTH, Include/cSIFT3D.h:27 DESC_NUMEL */
#define S3D_KP_BYTES 176    /* sizeof(CPUSIFT::Keypoint), Include/cSIFT3D.h:52-70 */

/* Keypoint record written by s3d_get_keypoints / s3d_get_extrema: byte-identical to
 * CPUSIFT::Keypoint (Include/cSIFT3D.h:52-70).  `desc` is written as NULL; the caller owns the
 * descriptor buffer and patches the pointer (the façade does). */
typedef struct s3d_keypoint {
    float x, y, z;
    float scale;
    int octave, level;
    float rx, ry, rz;
    float win[3];
    float eigvalue[3];
    float eigvector[9];
    float Rotation[9];
    float str_tensor[9];
    float* desc;
} s3d_keypoint;

/* Constructor parameters: the default arguments of CSIFT3DFactory::CreateCSIFT3D
 * (Include/cSIFT3D.h:187-202; defaults Include/cSIFT3D.h:13-21). */
typedef struct s3d_params {
    int num_kp_levels;     /* 3    */
    float sigma_default;   /* 1.6  */
    float sigma_n_default; /* 1.15 */
    float peak_thresh;     /* 0.1  */
    float max_eig_thres;   /* 0.9  */
    float corner_thresh;   /* 0.4  */
    int device;            /* CUDA device ordinal; -1 = current device */
    int keep_levels;       /* 1 = keep the pyramids after s3d_run (== building the reference with
                              CHECK_ENABLE, Src/cSIFT3D.cc:223-225) so s3d_get_level works */
    int exact_recheck;     /* 1 = re-evaluate orientation candidates whose accept/reject tests are
                              within a small margin in the reference's serial FP32 order */
    int profile;           /* 1 = bracket every kernel launch with CUDA events (s3d_get_kernel_stats) */
    void* stream;          /* cudaStream_t to run on (caller keeps it alive); NULL = a private stream;
                              (void*)1 = cudaStreamLegacy, the default stream */
} s3d_params;

typedef struct s3d_ctx* s3d_handle;

/* ---- library ------------------------------------------------------------------------------ */
S3D_API int s3d_version(void);
S3D_API const char* s3d_last_error(void);
/* Number of usable sm_100 devices (0 when none: every compute call will fail). */
S3D_API int s3d_device_count(void);
/* Checks on the device that FP32 a*b+c is NOT contracted and that division/sqrt are IEEE
 * (the bit-exact contract of the dense stages).  0 = ok. */
S3D_API int s3d_selftest(int device);
S3D_API void s3d_default_params(s3d_params* p);
/* Number of kernels this library has launched in this process (bench.py "gpu_launches"). */
S3D_API uint64_t s3d_launch_count(void);

/* ---- extraction --------------------------------------------------------------------------- */
/* CSIFT3DFactory::CreateCSIFT3D(float*, nx, ny, nz, ...)  Src/cSIFT3D.cc:103-110 → ctor :146-163
 * (copy the caller's volume, divide by global max|v|: data_scale Src/cUtil.cc:536-564).
 * `vol` is HOST memory, x fastest (xs=1, ys=nx, zs=nx*ny; Src/Util/cTexImage.cc:28-30); it is
 * copied and never written.  Pinned host memory is copied asynchronously. */
S3D_API int s3d_create(const float* vol, int nx, int ny, int nz, const s3d_params* p, s3d_handle* out);
/* Same, without waiting for the copy: the H2D transfer, max|v| and the division are only ENQUEUED on
 * the handle's stream.  `vol` must be pinned host memory and must stay alive and unchanged until
 * s3d_run / s3d_wait on this handle has returned.  Lets the upload of volume k+1 overlap the
 * extraction of volume k (each handle has its own stream unless params.stream is set). */
S3D_API int s3d_create_async(const float* vol, int nx, int ny, int nz, const s3d_params* p, s3d_handle* out);
/* Same, for a volume already resident in device memory (HBM-resident timing, chained pipelines).  d_vol is contiguous
 * float32 on params.device.  ORDERING: the volume is read on the handle's stream — params.stream when given, else a
 * private non-blocking stream that is NOT ordered against the stream that produced d_vol: either pass the producing
 * stream in params.stream or make sure the producer has finished.  The call returns after the normalised copy has been
 * made; d_vol is not needed afterwards.  (All entry points leave `params.device` as the thread's current device.) */
S3D_API int s3d_create_device(const float* d_vol, int nx, int ny, int nz, const s3d_params* p, s3d_handle* out);
/* CSIFT3D::KpSiftAlgorithm()  Src/cSIFT3D.cc:165-235: Initialize, Gaussian scale space, DoG,
 * detection, orientation, description.  Blocks until the results are on the host side of the
 * handle (keypoint records + descriptors). */
S3D_API int s3d_run(s3d_handle h);
/* Asynchronous halves of s3d_run: enqueue everything on the handle's stream / wait for it. */
S3D_API int s3d_run_async(s3d_handle h);
S3D_API int s3d_wait(s3d_handle h);
/* CSIFT3D::~CSIFT3D  Src/cSIFT3D.cc:140-144 */
S3D_API void s3d_destroy(s3d_handle h);

/* octave_num  Src/cSIFT3D.cc:254-255 */
S3D_API int s3d_num_octaves(s3d_handle h, int* n);
/* dims of an octave (n >> o, Src/cUtil.cc:219-221) */
S3D_API int s3d_level_dims(s3d_handle h, int octave, int* dims3);
/* CSIFT3D::GetKeypoints()  Src/cSIFT3D.cc:1686-1688: count, then records (+ K x 768 descriptors;
 * either pointer may be NULL).  Order = (octave, level, z, y, x) of the surviving detections. */
S3D_API int s3d_num_keypoints(s3d_handle h, int* n);
S3D_API int s3d_get_keypoints(s3d_handle h, s3d_keypoint* kp, float* desc);
/* The same copies ENQUEUED on the handle's stream without waiting (kp / desc should be pinned host memory, and must stay
 * valid until s3d_sync returns): lets the caller start the next volume's extraction on another handle while this one's
 * results travel to the host.  s3d_sync = wait for everything enqueued on the handle's stream (records the D2H time). */
S3D_API int s3d_get_keypoints_async(s3d_handle h, s3d_keypoint* kp, float* desc);
S3D_API int s3d_sync(s3d_handle h);
/* The handles' device buffers are recycled through a size-class cache inside the library (a freed pyramid serves the next
 * volume, whatever stream it runs on).  s3d_trim_cache returns every cached block of `device` (-1 = current) to the driver;
 * *cached_bytes (may be NULL) receives the number of bytes that were cached. */
S3D_API int s3d_trim_cache(int device, unsigned long long* cached_bytes);
/* The raw detections after orientation (`extre`, Src/cSIFT3D.cc:410,427-456): records carry
 * str_tensor / win / eigvalue / eigvector, rejected ones x=y=z=-1; codes[i] in {1,-1,-2,-3}
 * (RET, Src/cSIFT3D.cc:445).  xyz5 (may be NULL) receives the integer x,y,z,octave,level of every
 * detection — the bit-exact detection mask (level_extrema, Src/cSIFT3D.cc:412,419). */
S3D_API int s3d_num_extrema(s3d_handle h, int* n);
S3D_API int s3d_get_extrema(s3d_handle h, s3d_keypoint* kp, int* codes, int* xyz5);
/* GET_GSS() / GET_DOG()  Include/cSIFT3D.h:169-174.  which: 0 = Gaussian level idx (o*(L+3)+i),
 * 1 = DoG level idx (o*(L+2)+i).  Needs keep_levels. */
S3D_API int s3d_get_level(s3d_handle h, int which, int idx, float* out);
/* Level metadata as TexImage carries it (Include/Util/cTexImage.h:13-25): dims3 = nx,ny,nz;
 * meta4 = scale, ux, uy, uz (Src/cUtil.cc:209-224). */
S3D_API int s3d_level_info(s3d_handle h, int which, int idx, int* dims3, float* meta4);
/* Device address of the K x 768 descriptor block of a finished extraction (valid until
 * s3d_destroy): lets a matcher consume descriptors without the host round trip. */
S3D_API int s3d_device_descriptors(s3d_handle h, const float** d_desc, int* n);
/* Device pointers of every result array of a finished run (valid until s3d_destroy), for consumers
 * that gather / merge on the GPU (z-slab shards): d_ptrs5 = {keypoint records (n_kps x 176 B, desc
 * pointer null), descriptors (n_kps x 768 float), detection records (n_extre x 176 B), accept codes
 * (n_extre int), xyz5 (n_extre x 5 int: x, y, z, octave, level)}. */
S3D_API int s3d_device_results(s3d_handle h, const void** d_ptrs5, int* n_kps, int* n_extre);
/* The normalised input (Host_Im after data_scale, Src/cSIFT3D.cc:162). */
S3D_API int s3d_get_input(s3d_handle h, float* out);
/* Per-level detection thresholds peak_thresh*max|DoG| (Src/cSIFT3D.cc:384-385), o*L + (i-1). */
S3D_API int s3d_get_thresholds(s3d_handle h, float* out, int n);
/* SIFT_TimerPara m_timer  Include/Util/common.h:22-41, filled Src/cSIFT3D.cc:228-233:
 * t[0]=alloc t[1]=gss t[2]=dog t[3]=detect t[4]=orient t[5]=describe t[6]=release t[7]=total
 * (seconds, from CUDA events on the handle's stream; t[8]=h2d, t[9]=d2h). */
S3D_API int s3d_get_timers(s3d_handle h, double* t10);
/* Bookkeeping of the exactness safeguards of one run: out4[0] = detections re-evaluated in the
 * reference's serial FP32 order (orientation), [1] = accept/reject codes that changed, [2] = keypoints
 * whose descriptor the fixed-point kernel handed to the FP32 kernel, [3] = times the sparse stage was repeated
 * because the optimistically sized detection / keypoint buffers were too small (a step has no host round trip:
 * counts stay on the device until s3d_wait; S3D_CAP_EXTRE / S3D_CAP_KPS in the environment force the capacities). */
S3D_API int s3d_get_counters(s3d_handle h, int* out4);
/* Descriptor accumulation (Extract_Descriptor_Imp Src/cSIFT3D.cc:1152-1381 adds 24 weighted
 * contributions per voxel into the 768-bin histogram): 0 = fixed point with native shared-memory
 * integer atomics + FP32 redo of keypoints that exceed the overflow-safe cap (default), 1 = FP32
 * staged/ordered accumulation for every keypoint, 2 = as 0 with a deliberately small scale margin
 * (exercises the redo path).  All paths meet the descriptor tolerance (cosine >= 0.9999). */
S3D_API int s3d_set_describe_path(int path);

/* Per-kernel-class device times of the last s3d_run (needs params.profile = 1): for class c <
 * *n_classes, ms[c] = summed CUDA-event time, launches[c], alg_bytes[c] = algorithmic bytes moved
 * (DESIGN.md states the per-voxel figures).  Arrays hold up to `cap` entries. */
S3D_API int s3d_get_kernel_stats(s3d_handle h, int cap, int* n_classes, double* ms, long long* launches,
                                 double* alg_bytes);
S3D_API const char* s3d_kernel_class_name(int cls);

/* ---- one large volume in z-slabs over several GPUs (BASELINE.json configs[2]) ----
 * The reference processes one volume in one address space (Src/cSIFT3D.cc:165-235); its z pass
 * (Src/cSIFT3D.cc:615-617) is the axis this path shards.  Shard r OWNS octave-0 planes [own0, own1) (plane k of
 * octave o belongs to the owner of plane k*2^o) and keeps local level buffers for owned planes +- `halo`.
 * Every Gaussian level is produced on owned +- 1 planes (so the DoG neighbours of detection are local); before
 * level i the shards exchange the hw_i + 1 source planes next to their borders, and once an octave is complete the
 * planes of levels 1..L that the orientation / descriptor windows of the neighbours' keypoints reach.  Octaves too
 * thin to shard are gathered once (their seed level) and computed by every shard ("replicated"); detection,
 * orientation and description always run on the owned planes only.  Input max|v| and the per-level max|DoG| are
 * all-reduced(max).  Results (levels on owned planes, detections, keypoints, descriptors) are bit-identical to the
 * unsharded run; the per-(octave, level) concatenation of the shards' lists in shard order is the reference's order.
 * All of it is enqueued on the shards' streams from C: the only host waits are the two count read-backs of the
 * sparse stage and the result gather.
 *
 * Transports.  s3d_comm wraps NCCL for one process per GPU (the caller distributes the 128-byte unique id, e.g.
 * with torch.distributed / MPI / a file) — collective calls: every rank calls the same function with the same
 * volume geometry.  s3d_extract_multi is the single-process form: one host thread per shard, device-to-device /
 * peer copies between the shards' buffers (shards may share a device: "logical shards", the CI check).
 */
typedef struct s3d_comm* s3d_comm_t;
/* Owned octave-0 planes of shard `rank` of `world`: contiguous, balanced, even starts. */
S3D_API int s3d_slab_bounds(int nz, int world, int rank, int* own0, int* own1);
/* local/owned plane ranges {za, zb, p0, p1} of a SHARDED octave (no handle needed) */
S3D_API int s3d_slab_extent(int nz, int own0, int own1, const s3d_params* p, int octave, int* out4);
/* First octave that is replicated rather than sharded for this geometry (== octave count: none). */
S3D_API int s3d_slab_first_replicated(int nx, int ny, int nz, int world, const s3d_params* p, int* octave);
/* The transfers of one halo fill as shard `rank` sees them (host logic, no GPU): planes [k0, k1) of a level of
 * `octave` that `rank` receives from (kind 0) or sends to (kind 1) shard `peer`, for a halo of `depth` planes
 * beyond the owned range (depth < 0: every plane the shard does not own — the gather of a replicated octave).
 * Writes up to cap (peer, kind, k0, k1) quadruples, returns the count in *n. */
S3D_API int s3d_slab_plan(int nz, int world, int rank, int octave, int depth, int* quads, int cap, int* n);

S3D_API int s3d_comm_unique_id(void* id128);
S3D_API int s3d_comm_create(const void* id128, int world, int rank, int device, s3d_comm_t* out);
S3D_API void s3d_comm_destroy(s3d_comm_t comm);
S3D_API int s3d_comm_info(s3d_comm_t comm, int* world, int* rank, int* device);
/* Bytes this rank has sent / received through the communicator since its creation. */
S3D_API int s3d_comm_traffic(s3d_comm_t comm, unsigned long long* sent, unsigned long long* received);

/* CreateCSIFT3D + KpSiftAlgorithm of ONE volume over the ranks of `comm` (collective).  vol_own = this rank's
 * OWNED raw planes [own0, own1) of s3d_slab_bounds (host memory, pinned for an asynchronous upload, or device
 * memory with on_device != 0).  Blocks until this rank's part is done; *out holds the rank's own keypoints
 * (reference order within the shard) until s3d_slab_gather. */
S3D_API int s3d_slab_run(s3d_comm_t comm, const float* vol_own, int on_device, int nx, int ny, int nz,
                         const s3d_params* p, s3d_handle* out);
/* The two halves of s3d_slab_run.  s3d_slab_create (CreateCSIFT3D for this rank's shard) only ENQUEUES the
 * allocation, the copy of the owned planes and the first sweep of data_scale on the shard's stream and involves no
 * other rank: the next volume can upload while the current one is extracted (vol_own: pinned, alive until
 * s3d_slab_execute returns).  s3d_slab_execute (KpSiftAlgorithm) is the collective part and blocks.  Ranks must
 * execute their handles in the same order. */
S3D_API int s3d_slab_create(s3d_comm_t comm, const float* vol_own, int on_device, int nx, int ny, int nz,
                            const s3d_params* p, s3d_handle* out);
S3D_API int s3d_slab_execute(s3d_comm_t comm, s3d_handle h);
/* s3d_slab_execute without the final wait: kernels and collectives of the whole run are only enqueued (there is no
 * host round trip inside a run), s3d_wait(h) completes it.  Lets every rank enqueue volume k+1 while volume k runs. */
S3D_API int s3d_slab_execute_async(s3d_comm_t comm, s3d_handle h);
/* Collective: merge every rank's results in the reference's order into rank `root`'s handle (its
 * s3d_num_keypoints / s3d_get_keypoints / s3d_device_descriptors / s3d_get_extrema then answer for the whole
 * volume); with_extrema = 0 skips the per-detection debug records.  The other ranks keep their own part. */
S3D_API int s3d_slab_gather(s3d_comm_t comm, s3d_handle h, int root, int with_extrema);
/* Per-phase device time of this rank's last s3d_slab_run (ms): [0] all-reduce of max|v| + normalise, [1] pyramid
 * incl. halo exchanges, [2] window-halo exchange, [3] all-reduce + sparse stages, [4] gather (filled by
 * s3d_slab_gather), [5] allocation + copy of the owned planes + max|v| (s3d_slab_create), [6..7] reserved. */
S3D_API int s3d_slab_phases(s3d_handle h, double* ms8);

/* The same in ONE process: `nshards` shards, shard g on device devices[g] (repeats allowed; NULL = all on the
 * current device), `vol` = the whole volume in host memory.  handles[0] receives the merged results, handles[g]
 * shard g's own part (all must be destroyed by the caller; keep_levels keeps every shard's local levels alive
 * for s3d_slab_level_buffer / s3d_get_level).  This is what CSIFT3DFactory::CreateCSIFT3D + KpSiftAlgorithm do
 * when SIFT3D_B200_DEVICES names more than one device. */
S3D_API int s3d_extract_multi(const float* vol, int nx, int ny, int nz, const s3d_params* p, const int* devices,
                              int nshards, int with_extrema, s3d_handle* handles);
S3D_API int s3d_slab_info(s3d_handle h, int* noct, int* halo, int* levels_per_octave, int* first_replicated_octave);
/* Device address of the local planes of a level (which: 0 Gaussian, 1 DoG) and {za, zb, p0, p1}. */
S3D_API int s3d_slab_level_buffer(s3d_handle h, int which, int idx, float** d_ptr, int* ext4);

/* ---- free kernels (parity hooks ≙ Include/cSIFT3D.h:208-239) ------------------------------- */
/* GaussianSmooth_3D  Src/cSIFT3D.cc:535-622 (host buffers in/out). */
S3D_API int s3d_gaussian_smooth(const float* src, int nx, int ny, int nz, float sigma, float* dst);
/* One separable pass (GaussianSmooth_3D_Imp, Src/cSIFT3D.cc:624-790) along axis 0/1/2 with the
 * given taps; variant 0 = generic kernel, 1 = fast (vector / marching) kernel when eligible. */
S3D_API int s3d_blur_axis(const float* src, int nx, int ny, int nz, int axis, const float* w, int hw,
                          int variant, float* dst);
/* DownSample_3D  Src/cSIFT3D.cc:506-533 */
S3D_API int s3d_downsample(const float* src, int nx, int ny, int nz, float* dst);

/* ---- matching ------------------------------------------------------------------------------ */
/* muBruteMatcher::injectMatch / bijectMatch / enhancedMatch  Src/cMatcher.cc:218-228 →
 * bijectMatchBase :146-215.  type: 1 inject, 2 biject, 3 enhanced (Include/cMatcher.h:14-18).
 * ref_desc / tar_desc: HOST, row-major n x 768 (the façade gathers Keypoint::desc).
 * Outputs (HOST, caller-allocated, any may be NULL):
 *   gIdx,gDist,sIdx,sDist   length n_ref: glodenIdx (post-filter: rejected = negated,
 *                            Src/cMatcher.cc:92-94,141-142), glodenDistSquare, silverIdx,
 *                            silverDistSquare (Src/cMatcher.cc:71-77)
 *   gIdx2,gDist2,sIdx2,sDist2 length n_tar: the reverse (masked) search
 *   pair_ref,pair_tar       length n_ref: matched index pairs in ascending ref order
 *                            (toCvec, Src/cMatcher.cc:99-112); *n_pairs = count
 *   times3                  matchTime, revMatchTime, totalTime (seconds)
 */
S3D_API int s3d_match(int type, const float* ref_desc, int n_ref, const float* tar_desc, int n_tar, double thr,
                      int* gIdx, float* gDist, int* sIdx, float* sDist, int* gIdx2, float* gDist2, int* sIdx2,
                      float* sDist2, int* pair_ref, int* pair_tar, int* n_pairs, double* times3);
/* s3d_match where either descriptor set may already live in device memory (ref_on_device /
 * tar_on_device != 0); outputs are HOST arrays as in s3d_match. */
S3D_API int s3d_match_ex(int type, const float* ref_desc, int n_ref, int ref_on_device, const float* tar_desc, int n_tar,
                         int tar_on_device, double thr, int* gIdx, float* gDist, int* sIdx, float* sDist, int* gIdx2,
                         float* gDist2, int* sIdx2, float* sDist2, int* pair_ref, int* pair_tar, int* n_pairs,
                         double* times3);
/* Search path of calMatches: 0 = auto (tensor-core candidate pass + exact re-rank for large
 * searches, exact CUDA-core kernel for small ones), 1 = exact kernel only, 2 = tensor cores always
 * (kernel variant chosen by size), 3 = tensor cores with one CTA per tile, 4 = tensor cores with CTA
 * pairs (cta_group::2) and the query tile resident in shared memory.
 * Results are identical on every path (the tensor-core pass proves its candidate set complete or
 * falls back per row).  The proof needs descriptors as the reference produces them (non-negative, entries <= 1,
 * finite: Src/cSIFT3D.cc:1350-1358); the conversion pass checks this on the device and a set with any other entry
 * is searched by the exact kernel whatever the path says. */
S3D_API int s3d_set_match_path(int path);
/* Rows searched on the tensor-core path and rows that needed the exact fallback, since start/reset. */
S3D_API void s3d_match_stats(unsigned long long* tc_rows, unsigned long long* fallback_rows, int reset);
/* Same with DEVICE-resident descriptor sets and DEVICE outputs (HBM-resident timing; the
 * extract→match handoff of SURVEY.md §8f-1).  stream is a cudaStream_t (0 = default). */
S3D_API int s3d_match_device(int type, const float* d_ref, int n_ref, const float* d_tar, int n_tar, double thr,
                             int* d_gIdx, float* d_gDist, int* d_sIdx, float* d_sDist, int* d_gIdx2,
                             float* d_gDist2, int* d_sIdx2, float* d_sDist2, int* d_pair_ref, int* d_pair_tar,
                             int* d_n_pairs, void* stream);
/* calMatches (Src/cMatcher.cc:40-79) over ONE shard of the database: per query the best and
 * second-best (dot, global index) over db rows [0, n_db) reported as index + db_offset; dots are
 * the reference's double sums (KP_squareSum :17-23).  mask (may be NULL): queries with 0 are
 * skipped (idx = -1).  All pointers DEVICE.  Used for multi-GPU sharding. */
S3D_API int s3d_top2_device(const float* d_q, int n_q, const float* d_db, int n_db, int db_offset,
                            const int* d_mask, double* d_dot1, int* d_idx1, double* d_dot2, int* d_idx2,
                            void* stream);
/* Merge `parts` partial top-2 lists (each n_q long, laid out [part][n_q]) under the total order
 * (dot desc, index asc) — equal to the reference's sequential strict-'>' scan — and emit the
 * distances 2-2*dot and indices of calMatches (Src/cMatcher.cc:71-77).  DEVICE pointers. */
S3D_API int s3d_top2_merge_device(int parts, int n_q, const double* d_dot1, const int* d_idx1,
                                  const double* d_dot2, const int* d_idx2, const int* d_mask, float* d_gDist,
                                  int* d_gIdx, float* d_sDist, int* d_sIdx, void* stream);
/* filter / countMatched+toMask / bijectFilter / toCvec (Src/cMatcher.cc:81-144) on DEVICE arrays. */
S3D_API int s3d_ratio_filter_device(int* d_gIdx, const float* d_gDist, const float* d_sDist, int n, double thr,
                                    void* stream);
S3D_API int s3d_count_mask_device(const int* d_gIdx, int n_ref, int* d_mask, int n_tar, int count_thres,
                                  void* stream);
S3D_API int s3d_biject_filter_device(int* d_gIdx, int n_ref, const int* d_mask, const int* d_gIdx2, void* stream);
S3D_API int s3d_pairs_device(const int* d_gIdx, int n_ref, int* d_pair_ref, int* d_pair_tar, int* d_n_pairs,
                             void* stream);

/* muBruteMatcher over several GPUs (BASELINE.json configs[4]): the searched set of each
 * direction (the inner database loop of calMatches, Src/cMatcher.cc:58) is sharded into contiguous index ranges, one
 * per rank, for the tensor-core candidate pass; the approximate top-8 lists are re-distributed so that the exact
 * re-rank is sharded by query (every rank re-ranks 1/world of the queries against the full database); the exact
 * top-2 blocks are all-gathered and the filters run replicated.  Collective: every rank passes the FULL sets
 * (DEVICE memory, replicated) and receives the complete outputs (DEVICE memory).  Bit-identical to s3d_match_device. */
S3D_API int s3d_match_sharded(s3d_comm_t comm, int type, const float* d_ref, int n_ref, const float* d_tar, int n_tar,
                              double thr, int* d_gIdx, float* d_gDist, int* d_sIdx, float* d_sDist, int* d_gIdx2,
                              float* d_gDist2, int* d_sIdx2, float* d_sDist2, int* d_pair_ref, int* d_pair_tar,
                              int* d_n_pairs, void* stream);
/* The same in ONE process over ndev devices (one host thread per device, peer copies); HOST sets in, HOST outputs
 * out as s3d_match.  devices == NULL: ndev logical shards on the current device.  This is what muBruteMatcher does
 * when SIFT3D_B200_DEVICES names more than one device. */
S3D_API int s3d_match_multi(int type, const float* ref_desc, int n_ref, const float* tar_desc, int n_tar, double thr,
                            const int* devices, int ndev, int* gIdx, float* gDist, int* sIdx, float* sDist, int* gIdx2,
                            float* gDist2, int* sIdx2, float* sDist2, int* pair_ref, int* pair_tar, int* n_pairs,
                            double* times3);

/* ---- volume ingest --------------------------------------------------------- */
/* readNiiFile (Include/Util/readNii.h:6, Src/Util/readNii.cpp:5-39): single-file NIfTI-1/-2, plain
 * or gzip, any scalar datatype cast to float32 as stored (no scl_slope/scl_inter, like the
 * reference).  Returns a host buffer of nx*ny*nz floats (x fastest) to release with s3d_free_host,
 * or NULL. */
S3D_API float* s3d_read_nii(const char* path, int* nx, int* ny, int* nz);
S3D_API void s3d_free_host(float* p);

#ifdef __cplusplus
}
#endif
#endif /* SIFT3D_B200_H */
