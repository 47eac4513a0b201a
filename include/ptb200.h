/* ptb200.h -- C ABI of libptb200.so, the B200 (sm_100a) drop-in for the per-pixel radiance operator of
 * KVM-Explorer/AscendPathTracing.  Citations are file:line into the reference repository.
 *
 * Plain pointers and sizes only; no C++/torch types.  Unless a function says "host", every buffer
 * pointer is a DEVICE pointer owned by the caller, and every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*, NULL = the legacy default stream) exactly like the reference's
 * render_do (src/render.cpp:264-266; the caller synchronises, src/main.cpp:75).
 *
 * Threads and streams: every entry may be called from several host threads at once (the last-error text is per thread; the
 * library's workspace arenas hand out blocks under one lock and grow by adding an arena while blocks are held).  `stream`
 * must belong to the CURRENT device (checked: cudaErrorInvalidDevice otherwise).  On one device the trace launches of
 * different streams run ONE AFTER THE OTHER, not side by side: the staged scene (constant bank), the ray-generator block
 * and the chunk dispenser of the persistent kernels are per-device singletons, and a launch sequence waits on an event for
 * the previous one to finish reading them.  Streams therefore overlap copies and resolves with a trace, never two traces --
 * one trace launch already fills all 148 SMs.  Different devices are fully independent.
 *
 * Return values: 0 on success, otherwise a negative PTB200_E* code or a positive cudaError_t.
 * ptb200_last_error() returns a thread-local human-readable message for the last failure.
 * There is no CPU fallback anywhere: without a CUDA device every compute entry fails with
 * PTB200_ENODEV (or the CUDA error).
 *
 * Buffer layouts (identical to the reference's files):
 *   rays     float32 SoA [6][N]   ox,oy,oz,dx,dy,dz planes           (scripts/gen_data.py:65-71)
 *   spheres  float32 SoA [10][stride] r^2,x,y,z,ex,ey,ez,cr,cg,cb    (scripts/gen_data.py:106-127,
 *            src/rt_helper.h:91-103); the reference is stride = 8 padded to 512 bytes
 *   colors   float32 SoA [3][N]   R,G,B planes                        (src/render.cpp:218-220)
 *   image    uint8 [H rows, top first][W][3]                          (scripts/data_visualization.py:11-57)
 *   N = width * height * 4 * samples, path index ((((x*H + y)*2 + sy)*2 + sx)*S + k)  (gen_data.py:32-36)
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB200_ABI_VERSION 2

/* exported from libptb200.so (the library is built with -fvisibility=hidden) */
#if defined(__GNUC__)
#define PTB200_API __attribute__((visibility("default")))
#else
#define PTB200_API
#endif

enum {
    PTB200_OK = 0,
    PTB200_EINVAL = -1,  /* bad argument (null pointer, size rule violated, ...) */
    PTB200_ENODEV = -2,  /* no CUDA device */
    PTB200_ENOMEM = -3,  /* arena exhausted / allocation failed */
    PTB200_EIO = -4      /* file I/O failed */
};

/* Run-time form of the reference's compile-time configuration (src/common.h:4-14) and literals:
 * depth 5 (src/render.cpp:141), light index 7 (src/rt_helper.h:776), emission 12 (src/render.cpp:194-196). */
typedef struct PtParams {
    int32_t width;          /* WIDTH   (common.h:4)  */
    int32_t height;         /* HEIGHT  (common.h:5)  */
    int32_t samples;        /* SAMPLES (common.h:6): samples per sub-pixel; spp = 4*samples */
    int32_t depth;          /* bounces per path (render.cpp:141), reference = 5 */
    int32_t sphere_count;   /* SPHERE_NUM (common.h:10), reference = 8 */
    int32_t sphere_stride;  /* floats between members in the sphere SoA; reference = 8 */
    int32_t light_index;    /* sphere that ends a path (rt_helper.h:776), reference = 7 */
    float emission_scale;   /* final multiplier (render.cpp:194-196), reference = 12 */
    int32_t flags;          /* PTB200_F_* */
    int32_t column_step;    /* ptb200_render_image*: 0 or 1 = every column of [x0, x1); k > 1 = columns x0, x0+k, x0+2k, ... < x1
                             * into a dense image of ceil((x1-x0)/k) columns (multi-GPU: rank r of G renders x0 = r, k = G in
                             * ONE launch, evenly loaded whatever the scene).  Ignored by the other entry points. */
} PtParams;

enum {
    /* Stop tracing a path once its colour can no longer change (light hit, or throughput exactly 0).
     * Bit-identical results to the reference's fixed-depth loop. Default on;
     * set PTB200_F_FIXED_DEPTH to trace every bounce like the reference does. */
    PTB200_F_FIXED_DEPTH = 1
};

/* Fills *p with the reference defaults (16x16, SAMPLES=1, depth 5, 8 spheres, light 7, x12). */
PTB200_API void ptb200_default_params(PtParams *p);

PTB200_API int ptb200_abi_version(void);
PTB200_API const char *ptb200_last_error(void);
/* Number of CUDA devices visible (0 when none / no driver). Never fails. */
PTB200_API int ptb200_device_count(void);

/* ---- the reference's kernel entry points ------------------------------------------------------ */

/* Replaces `extern "C" __global__ __aicore__ void render(GM_ADDR rays, GM_ADDR spheres, GM_ADDR colors)`
 * (src/render.cpp:253-259, declared src/main.cpp:13).  Same three opaque byte pointers in the same
 * order.  Sizes come from the library's legacy configuration (ptb200_set_legacy_config; initially the
 * reference's common.h values).  Launches on the default stream and returns after completion, like
 * ICPU_RUN_KF (src/main.cpp:37).  Errors are logged to stderr in the style of CHECK_ACL
 * (src/data_utils.h:41-47) because the signature returns void. */
PTB200_API void render(uint8_t *rays, uint8_t *spheres, uint8_t *colors);

/* Replaces `void render_do(uint32_t blockDim, void *l2ctrl, void *stream, uint8_t *rays, uint8_t *spheres,
 * uint8_t *colors)` (src/render.cpp:264-266, declared src/main.cpp:9-10).  Asynchronous on `stream`.
 * blockDim and l2ctrl are accepted and ignored (the reference passes 8 and nullptr, src/main.cpp:18,74). */
PTB200_API void render_do(uint32_t blockDim, void *l2ctrl, void *stream, uint8_t *rays, uint8_t *spheres, uint8_t *colors);

/* The reference fixes W/H/SAMPLES at compile time (src/common.h:4-6, src/render.cpp:256); the two legacy
 * entry points above use this process-wide configuration instead. Validates the reference's preconditions
 * (src/render.cpp:68-73: N % 8 == 0 and (N/8) % 128 == 0). */
PTB200_API int ptb200_set_legacy_config(const PtParams *p);
PTB200_API void ptb200_get_legacy_config(PtParams *p);

/* Run-time-parameter sibling of render_do.  Traces paths [first, first+count) of the N-path buffers
 * (count < 0 means "to the end"); the reference's per-core slice (src/render.cpp:24-27) is first =
 * b*N/8, count = N/8.  Any N >= 0 is accepted (no divisibility rule). */
PTB200_API int render_do_ex(const PtParams *p, void *stream, const uint8_t *rays, const uint8_t *spheres, uint8_t *colors,
                 int64_t first, int64_t count);

/* ---- the steps either side of the kernel ---------------------------------------- */

/* Camera rays of image columns [x0, x1) (scripts/gen_data.py:21-75) written as float32 SoA [6][M],
 * M = (x1-x0)*H*4*S, in the reference's order.  `uniforms` (device, 2 doubles per ray) replays a
 * recorded random stream -- the reference's is NumPy MT19937 seed 0, see ptb200_mt19937_uniforms --
 * and may be NULL, in which case the counter-based generator (Philox4x32-10 keyed by `seed`, counter =
 * global path index) supplies them. */
PTB200_API int ptb200_gen_rays(const PtParams *p, void *stream, const double *uniforms, uint64_t seed, int32_t x0, int32_t x1,
                    float *rays);

/* HOST helper: the doubles np.random.seed(seed); np.random.rand() yields (gen_data.py:37-39,438),
 * skipping the first `skip` doubles; ray i consumes doubles 2i and 2i+1.  Writes n doubles to host memory. */
PTB200_API int ptb200_mt19937_uniforms(uint32_t seed, uint64_t skip, uint64_t n, double *out_host);

/* HOST helper: the 512-byte scene of scripts/gen_data.py:92-132 (128 floats) into host memory. */
PTB200_API int ptb200_default_scene(float *out128_host);

/* Resolve (scripts/data_visualization.py:20-59): colors float32 SoA [3][N] -> uint8 image of columns
 * [x0, x1): out is [H][x1-x0][3].  Bit-exact with NumPy's float32 pairwise mean. */
PTB200_API int ptb200_resolve(const PtParams *p, void *stream, const float *colors, int32_t x0, int32_t x1, uint8_t *image);

/* Fused production path: generate, trace and resolve columns [x0, x1) without materialising rays or
 * per-path colours.  uniforms as in ptb200_gen_rays (NULL -> counter-based RNG). image: [H][x1-x0][3].
 * `stats` (device, may be NULL) receives 2 uint64: paths traced, ray segments actually traced. */
PTB200_API int ptb200_render_image(const PtParams *p, void *stream, const uint8_t *spheres, const double *uniforms, uint64_t seed,
                        int32_t x0, int32_t x1, uint8_t *image, uint64_t *stats);

/* ---- material extension (NOT in the reference, parity against this repo's own CPU twin) ---- */

/* DIFF / SPEC / REFR bounce sampling with Russian roulette, "smallpt in binary32" (DESIGN.md section 8).
 * spheres: float32 SoA [11][stride] = the reference's ten rows + material (0 DIFF, 1 SPEC, 2 REFR); emission rows are used. */
typedef struct PtMaterialParams {
    int32_t max_depth;   /* hard cap on bounces per path */
    int32_t rr_start;    /* Russian roulette from this depth on (smallpt: 5) */
    float hit_epsilon;   /* self-intersection guard; 0.1 for binary32 with 1e5-radius walls (1e-4 leaks) */
    int32_t reserved;
    uint64_t seed;       /* Philox4x32-10 key; counter = (path index, bounce) */
} PtMaterialParams;

/* Fills *mp with max_depth 64, rr_start 5, hit_epsilon 0.1, seed 0. */
PTB200_API void ptb200_default_material_params(PtMaterialParams *mp);

/* Traces paths [first, first+count) of the N-path buffers with materials; element i uses RNG path index
 * path0 + (i - first) (pass the global index of `first` so that stripes and tiles reproduce the whole frame).
 * stats (device, nullable): number of ray segments traced is ADDED to stats[0]. */
PTB200_API int render_do_mat(const PtParams *p, const PtMaterialParams *mp, void *stream, const uint8_t *rays, const uint8_t *spheres,
                  uint8_t *colors, int64_t first, int64_t count, uint64_t path0, uint64_t *stats);

/* HOST helper: smallpt's 9-sphere scene (quoted at scripts/gen_data.py:77-89) as SoA [11][16] = 176 floats. */
PTB200_API int ptb200_smallpt_scene(float *out176_host);

/* Production composition with materials: generate (counter-based RNG, key = cam_seed), trace, resolve columns [x0, x1).
 * gamma != 0 applies smallpt's display transform pow(clamp(x), 1/2.2)*255 + 0.5 instead of the reference's
 * truncating, gamma-free quantiser. stats as in ptb200_render_image. */
PTB200_API int ptb200_render_image_mat(const PtParams *p, const PtMaterialParams *mp, void *stream, const uint8_t *spheres, uint64_t cam_seed,
                            int32_t x0, int32_t x1, int32_t gamma, uint8_t *image, uint64_t *stats);

/* ---- large scenes: GPU-built sphere BVH (BASELINE config C4; NOT in the reference) ------- */

typedef struct PtBvh PtBvh;
/* Builds an LBVH (Morton order, Karras hierarchy) over the spheres of an 11-row SoA [11][stride] on the device.
 * Spheres of radius >= 100 (the 1e5-radius walls, the light) stay in a brute-force list.  The nearest hit through the tree
 * equals the brute-force loop's bit for bit (t, index, lowest index on ties): every ray widens the boxes by the margin the
 * reference's binary32 test needs at its distance and direction length, and rays no margin can cover (origins 64 scene
 * widths away, directions more than 1 % off unit length, NaNs) are tested against every sphere.
 * Synchronous; the handle owns its device memory and keeps no reference to `spheres`. */
PTB200_API int ptb200_bvh_build(const uint8_t *spheres, int32_t count, int32_t stride, void *stream, PtBvh **out);
PTB200_API int ptb200_bvh_destroy(PtBvh *bvh);
PTB200_API int ptb200_bvh_info(const PtBvh *bvh, int32_t *n_spheres, int32_t *n_big, int32_t *n_small, int32_t *n_nodes);
/* Nearest hit only (the stage the tree replaces): rays SoA [6][n] -> tmin[n], index[n]; (1e20, 0) when nothing is hit. */
PTB200_API int ptb200_bvh_first_hit(const PtBvh *bvh, void *stream, const float *rays, int64_t n, float eps, float *tmin_out, int32_t *idx_out);
/* render_do_mat / ptb200_render_image_mat with the scene taken from the tree (p->sphere_count / sphere_stride are ignored). */
PTB200_API int render_do_mat_bvh(const PtParams *p, const PtMaterialParams *mp, const PtBvh *bvh, void *stream, const uint8_t *rays, uint8_t *colors,
                      int64_t first, int64_t count, uint64_t path0, uint64_t *stats);
PTB200_API int ptb200_render_image_mat_bvh(const PtParams *p, const PtMaterialParams *mp, const PtBvh *bvh, void *stream, uint64_t cam_seed, int32_t x0,
                                int32_t x1, int32_t gamma, uint8_t *image, uint64_t *stats);
/* HOST helper: the synthetic scene of BASELINE config C4: the reference's six walls and light
 * (scripts/gen_data.py:94-102, walls DIFF) followed by n_random spheres with centres uniform in [1,99]x[0,81.6]x[0,170],
 * radius uniform in [0.2,1], material uniform in {DIFF,SPEC,REFR}, colour uniform in [0.2,0.95]^3, drawn in that order
 * from NumPy-legacy MT19937(seed) doubles.  out: float32 SoA [11][stride], stride >= 7 + n_random. */
PTB200_API int ptb200_random_scene(int32_t n_random, uint32_t seed, int32_t stride, float *out_host);

/* ---- whole-job host-buffer entry (what the reference's main() does, src/main.cpp:46-92) --------- */

/* HOST buffers in, HOST buffer out: arena allocation, H2D of rays+spheres, render, D2H of colours,
 * synchronous.  This is the end-to-end call bench.py times. */
PTB200_API int ptb200_render_host(const PtParams *p, const float *rays_host, const float *spheres_host, float *colors_host);

/* ---- one process, several GPUs (the reference's 8-way split, src/render.cpp:9-10,24-27) -------------- */

/* The reference's kernel owns an 8-way split of the flat path array (blockDim = 8 AI cores, contiguous slices); its host
 * owns device selection, stream, copies, launch and synchronisation (src/main.cpp:46-92).  The two entries below are that
 * host for n_devices GPUs of one box: one host thread and one stream per device, the scene replicated, no exchange while
 * rendering.  devices == NULL means 0 .. n_devices-1; 1 <= n_devices <= 16; the caller's current device is restored.
 * A device may be listed more than once (its shares then run one after the other on it).
 *
 * ptb200_render_host_multi: ptb200_render_host with the N paths cut into n_devices contiguous slices (device r gets
 * [r*N/n, (r+1)*N/n), exactly the reference's per-core rule); every device streams its slice of the HOST ray planes in and
 * its slice of the HOST colour planes out.  Pinned host buffers (cudaMallocHost / cudaHostRegister) make the copies overlap.
 * ms_host (nullable, 1 + n_devices doubles): wall milliseconds of the call, then of every device's slice. */
PTB200_API int ptb200_render_host_multi(const PtParams *p, const int32_t *devices, int32_t n_devices, const float *rays_host,
                                        const float *spheres_host, float *colors_host, double *ms_host);

/* ptb200_render_image_multi: the production composition (generate, trace, resolve) of the whole W x H frame on n_devices
 * GPUs.  Device r renders image columns r, r + n, r + 2n, ... in ONE launch sequence (PtParams.column_step, which this entry
 * sets itself), so every GPU sees the same mix of cheap and expensive columns; the 8-bit column sets are then gathered on
 * devices[0] over NVLink (cudaMemcpyPeerAsync, peer access enabled once per pair), interleaved into the [H][W][3] frame by a
 * small kernel there and copied to `image`, which may be HOST memory or memory of devices[0].  Random numbers are keyed by
 * the global path index, so the frame is bit-identical for every n_devices (tests).
 *   mp == NULL: the reference-parity mirror kernel (spheres_host = SoA [10][stride]); otherwise the material extension
 *   (SoA [11][stride]); use_bvh != 0 (needs mp) builds the sphere BVH on every device first (scenes beyond 1024 spheres).
 *   seed: counter-based camera RNG key.  gamma as in ptb200_render_image_mat.
 *   stats_host (nullable, 2 uint64): paths traced, ray segments traced, summed over the devices.
 *   ms_host (nullable, 1 + n_devices doubles): wall milliseconds of the call (scene upload to image in place), then each
 *   device's own share (upload, BVH build, render, peer copy). */
PTB200_API int ptb200_render_image_multi(const PtParams *p, const PtMaterialParams *mp, int32_t use_bvh, int32_t gamma, const int32_t *devices,
                                         int32_t n_devices, const float *spheres_host, uint64_t seed, uint8_t *image, uint64_t *stats_host,
                                         double *ms_host);

/* ---- scene files beyond the reference's 512 bytes ------------------------------------------------ */

/* HOST helper: the layout of an input/spheres.bin of `bytes` bytes.
 *   512 bytes            the reference's file (src/main.cpp:24, scripts/gen_data.py:120-127): SoA [10][8] + 48 zero floats;
 *                        count = stride = 8 (SPHERE_NUM, src/common.h:10).  Read as an 11-row scene its material row is the
 *                        zero padding: every sphere DIFF.
 *   44 * stride bytes    SoA [11][stride]: the reference's ten rows (src/rt_helper.h:91-103) + material (0 DIFF, 1 SPEC,
 *                        2 REFR); count = stride minus trailing columns whose r^2 is 0 (padding), at least 1.
 * Anything else is PTB200_EINVAL.  `scene_host` may be NULL when only the stride is wanted (count is then = stride). */
PTB200_API int ptb200_scene_layout(const float *scene_host, size_t bytes, int32_t *count, int32_t *stride, int32_t *rows);

/* ---- device arena: replaces src/allocator.h's MemoryPool for the big buffers -------------------- */

typedef struct PtArena PtArena;
/* One cudaMalloc of `bytes`; first-fit free list with split on alloc and coalescing on free, the
 * semantics of Allocator::Init/Alloc/Free (src/allocator.h:54-151), 256-byte granularity. */
PTB200_API int ptb200_arena_create(size_t bytes, PtArena **out);
/* Same bookkeeping over memory the caller already owns (e.g. a framework tensor): [base, base+bytes) is
 * never dereferenced by the arena itself and is not freed by ptb200_arena_destroy. */
PTB200_API int ptb200_arena_wrap(void *base, size_t bytes, PtArena **out);
PTB200_API int ptb200_arena_destroy(PtArena *a);
/* Returns a device pointer or NULL when no free block is large enough (allocator.h:103-105 throws). */
PTB200_API void *ptb200_arena_alloc(PtArena *a, size_t bytes);
/* PTB200_EINVAL for a pointer the arena does not own or a double free (allocator.h:262-270 warns). */
PTB200_API int ptb200_arena_free(PtArena *a, void *ptr);
PTB200_API size_t ptb200_arena_capacity(const PtArena *a);
PTB200_API size_t ptb200_arena_in_use(const PtArena *a);
PTB200_API size_t ptb200_arena_largest_free(const PtArena *a);

/* ---- file I/O with the reference's semantics (src/data_utils.h:55-122), C linkage ---------------- */

/* Reads the whole file into buffer (fails if missing, empty or larger than buffer_size). */
PTB200_API int ptb200_read_file(const char *path, size_t *file_size, void *buffer_host, size_t buffer_size);
/* O_RDWR|O_CREAT|O_TRUNC, mode 0600 (data_utils.h:108). */
PTB200_API int ptb200_write_file(const char *path, const void *buffer_host, size_t size);
/* ASCII P3 writer (scripts/data_visualization.py:11-17); image is host uint8 [H][W][3]. */
PTB200_API int ptb200_write_ppm(const char *path, int32_t width, int32_t height, const uint8_t *image_host);

/* ---- measurement support ----------------------------------------------------------------------- */

/* Measures FP32 issue throughput on the current device with dependent-free register-only kernels.
 * kind: 0 FFMA, 1 FADD/FMUL alternating, 2 FFMA2 (packed f32x2), 3 FADD2/FMUL2 alternating,
 *       4 FADD + FSETP/FSEL mix, 5 rsqrtf, 6 IEEE sqrt (__fsqrt_rn), 7 IEEE div (__fdiv_rn), 8 FMUL2, 9 FADD2,
 *       10 FMUL2 + 2 scalar FADD (the exact kernel's pattern), 11 FADD2 + 2x(FSETP+FSEL), 12 MUFU.RSQ, 13 FMNMX,
 *       14 FFMA2 + LOP3, 15 FFMA2 with three register sources, 16 FFMA2 + FADD + LOP3, 17 FMUL2 with two register sources.
 * Writes giga-operations per second (one packed op counts as 2 lane-ops; FFMA counts as 1 op here --
 * multiply by 2 for FLOP) and the kernel milliseconds. */
PTB200_API int ptb200_measure_fp32(int32_t kind, int32_t iters, double *gops_out, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
