// Radiance kernel for sm_100a: the B200-native replacement of the reference's KernelRender
// (src/render.cpp:18-223) and its helpers (src/rt_helper.h).
//
// Mapping (not a port): the reference tiles 64 rays through ~30 vector ops per sphere with a scratch
// allocator in between; here one CUDA lane owns one path at a time and keeps its whole state in
// registers.  The kernel is persistent: the grid is sized to what the 148 SMs hold at once, and every
// lane walks its own strided sequence of paths (i, i + lanes, i + 2*lanes, ...).  The loop body is exactly
// one bounce, so lanes of a warp may be at different bounces of different paths without any
// divergence: a lane whose path has settled (light reached / throughput zero -- bit-identical early
// termination) swaps in its next path, already prefetched, while its neighbours carry on.
// The scene lives in the constant bank (immediate operands) plus a shared copy for per-lane lookups.
#include <cuda_pipeline.h>

#include <cstdlib>

#include "pt_device.cuh"
#include "pt_host.h"
#include "pt_material.cuh"
#include "pt_raygen.cuh"

namespace ptb200 {

// ---- scene staging: device SoA [10][stride] -> constant bank -------------------------------------
// Writes through the global-memory alias of the __constant__ symbols (cudaGetSymbolAddress); the
// constant cache is coherent across kernel launches, and launches on one stream are ordered.
__global__ void pack_scene_kernel(const float *__restrict__ spheres, int nsph, int stride, SceneConst *dst, int *zero_ok, float *zero_or_nan) {
    __shared__ int ok;
    if (threadIdx.x == 0)
        ok = 1;
    __syncthreads();
    const int padded = (nsph + 1) & ~1;  // the pairwise tests read one sphere past an odd count
    for (int k = threadIdx.x; k < padded; k += blockDim.x) {
        if (k >= nsph) {  // never-hit sphere: NaN centre -> NaN roots -> both compares false
            dst->nr2[k] = 0.0f;
            dst->cx[k] = dst->cy[k] = dst->cz[k] = __int_as_float(0x7fc00000);
            continue;
        }
        dst->nr2[k] = -spheres[0 * stride + k];
        dst->cx[k] = spheres[1 * stride + k];
        dst->cy[k] = spheres[2 * stride + k];
        dst->cz[k] = spheres[3 * stride + k];
        const float r = spheres[7 * stride + k], g = spheres[8 * stride + k], b = spheres[9 * stride + k];
        // zero-throughput early stop is exact only for finite colours with a clear sign bit
        const unsigned ur = __float_as_uint(r), ug = __float_as_uint(g), ub = __float_as_uint(b);
        if ((ur | ug | ub) >> 31 || !isfinite(r) || !isfinite(g) || !isfinite(b))
            ok = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *zero_ok = ok;
        *zero_or_nan = ok ? 0.0f : __int_as_float(0x7fc00000);
    }
}

// ---- the trace kernel -----------------------------------------------------------------------------
// Plane base pointers (already offset to the first path of the launch) travel as kernel parameters: the
// constant bank feeds IMAD.WIDE directly, so a lane forms an address with one instruction from its 32-bit
// path index instead of a 64-bit add chain.
struct TracePlanes {
    const float *ray[6];
    float *col[3];
};

constexpr int kRingBatches = 4;                  // batches of 32 rays in the ring per warp (power of two)
constexpr int kRing = 32 * kRingBatches;         // ring entries per warp
// Ring layout.  Fused generation (GEN): one entry = one ray as 8 floats (ox oy oz dx | dy dz - -): the generator stores it with
// one STS.128 + one STS.64 and a lane picks it up with one LDS.128 + one LDS.64.  Ray files: six planes of kRing floats, because
// the ring is filled by 4-byte cp.async copies from six global planes and entry-major rows would put the 32 lanes of a copy
// on 8 banks (measured: entry-major costs the ray-file kernel 1.4 %, and saves the fused kernel 3 %).  Same footprint either way.
constexpr int kRingEntry = 8;
constexpr int kRingFloats = kRing * kRingEntry;   // per warp (the plane-major layout uses 6/8 of it)
constexpr int kWarpsPerBlock = kTraceThreads / 32;

// A warp claims this many batches of 32 consecutive paths at a time (>= kRingBatches).  The last chunk a warp claims is
// the kernel's tail: with 64 batches (2048 paths) the trace kernel spent 5 % of its 2.1 ms, and the depth-50 sweep 10 %,
// waiting for the unluckiest warps; 8 batches cost one atomic per 256 paths and measure within 0.5 % of the best
// (tools/ab_chunks.sh, profiles/r1_chunk_size_ab.md).
#ifndef PTB_CHUNK_BATCHES
#define PTB_CHUNK_BATCHES 8
#endif
constexpr int kChunkBatches = PTB_CHUNK_BATCHES;

// Streams paths to one warp: chunks of 32 * kChunkBatches consecutive paths are claimed from a global counter (dynamic balancing:
// with a static split the slowest warp set the kernel's duration and a third of the warp slots sat idle at the end),
// fetched by coalesced cp.async batches of 32 rays into a shared-memory ring kRingBatches batches ahead of use, and
// handed to the lanes that ask for a path in ballot-rank order, i.e. consecutive indices to the lanes of one swap.
// Out of line on purpose: the binary64 generator needs ~40 registers of its own and runs once per 32 paths; as a call it
// borrows them for a moment instead of raising the register count (and lowering the occupancy) of the whole kernel.
// The generator's parameters live in the constant bank (staged per launch next to the scene, same stream ordering).
static __constant__ RayGenSource c_gen;

static __device__ __noinline__ void generate_ray_to_ring(unsigned int path, float *slot) {
    float r[6];
    generate_ray(c_gen, static_cast<long long>(path), r);
    *reinterpret_cast<float4 *>(slot) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float2 *>(slot + 4) = make_float2(r[4], r[5]);
}

// ---- fused resolve (FUSE): the production entries never materialise per-path colours ------------------------------------
// NumPy averages the S samples of a sub-pixel run in a fixed order (resolve_kernels.cu), and paths finish out of order here,
// so a finished path parks its colour in a small per-warp scratch -- [3 planes][2 chunk slots][kChunkPaths] floats of global
// memory that the warp keeps rewriting, i.e. that lives in L2 and costs no shared memory -- and when every path of a chunk
// has retired the warp itself reduces the chunk's runs in NumPy's order (128-bit L2 loads, two lanes per run block as in
// resolve_tiles_kernel) and stores ONE float per run and channel: the run's mean.  12 bytes per S paths leave the SM instead of
// 12 bytes per path, and the frame needs no colour workspace (C3: 0.4 GB of run means instead of 102 GB of colours written
// and read back).  Needs chunks made of whole runs whose NumPy recursion stays inside one chunk: S a power of two, 8..256.
constexpr int kChunkPaths = 32 * kChunkBatches;
constexpr int kFuseScratchFloats = 3 * 2 * kChunkPaths;  // per warp: [3 planes][2 chunk slots][kChunkPaths]
static_assert(kChunkPaths == 256, "the fused resolve assumes 256-path chunks (S up to 256 = NumPy's 128 + 128 split)");

struct FuseOut {
    float *mean[3];   // run means of this launch: plane c, element = (path of the launch) / S
    float *scratch;   // kFuseScratchFloats floats per warp of the grid
    unsigned int log2_s;
    float inv_s;      // 1 / S, a power of two: fl(sum * inv_s) == fl(sum / S), NumPy's float32 division
};
static __constant__ FuseOut c_fuse;

__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }

// Reduces the `valid` (a multiple of S) retired paths of one chunk slot: for every run and channel NumPy's pairwise sum
// (scripts/data_visualization.py:39-45 -> np.mean over a contiguous float32 axis) and the float32 division by S.
// Two lanes own one block of min(S, 128) samples: lane `half` holds accumulators r[4 half .. 4 half + 3] of NumPy's eight,
// fed in sample order by 128-bit loads; ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) is two in-lane adds and one shuffle; S = 256
// adds its two block sums (neighbouring lane pairs).  LG = log2(S) at compile time: the block loop unrolls completely.
template <int LG> __device__ __forceinline__ void fuse_reduce_chunk_s(const float *slot, float *const (&mean)[3], unsigned int first,
                                                                      unsigned int valid, unsigned int lane, float inv_s) {
    constexpr unsigned int kLgBlk = LG > 7 ? 7 : LG;  // NumPy's unrolled block: up to 128 samples
    constexpr unsigned int kBlk = 1u << kLgBlk;
    constexpr bool kTwoBlocks = LG > 7;               // S = 256: two blocks per run
    const unsigned int per_plane = valid >> kLgBlk;   // block tasks per colour plane
    const unsigned int total = 3u * per_plane;
    const unsigned int half = lane & 1u;
    const unsigned int run0 = first >> LG;
    for (unsigned int t0 = 0; t0 < total; t0 += 16u) {  // warp-uniform
        const unsigned int t = t0 + (lane >> 1);
        const bool live = t < total;
        const unsigned int tt = live ? t : 0u;           // idle pairs redo task 0 (they take part in the shuffles)
        const unsigned int ch = (tt >= per_plane ? 1u : 0u) + (tt >= 2u * per_plane ? 1u : 0u);
        const unsigned int rest = tt - ch * per_plane;   // block index inside the plane
        const float *a = slot + (ch * (2 * kChunkPaths) + (rest << kLgBlk) + 4u * half);
        constexpr unsigned int kLoads = kBlk / 8, kInFlight = kLoads < 4 ? kLoads : 4;  // L2 latency is paid kLoads / 4 times
        float4 acc = ldcg4(a);
#pragma unroll
        for (unsigned int g = 0; g < kLoads; g += kInFlight) {
            float4 v[kInFlight];
#pragma unroll
            for (unsigned int i = (g == 0 ? 1 : 0); i < kInFlight; i++)
                v[i] = ldcg4(a + 8u * (g + i));
#pragma unroll
            for (unsigned int i = (g == 0 ? 1 : 0); i < kInFlight; i++)
                acc.x = __fadd_rn(acc.x, v[i].x), acc.y = __fadd_rn(acc.y, v[i].y), acc.z = __fadd_rn(acc.z, v[i].z), acc.w = __fadd_rn(acc.w, v[i].w);
        }
        const float q = __fadd_rn(__fadd_rn(acc.x, acc.y), __fadd_rn(acc.z, acc.w));
        float r = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, 1));  // float addition commutes: both lanes hold the block sum
        if (kTwoBlocks)
            r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));    // the run's other block (neighbouring lane pair)
        const unsigned int run = kTwoBlocks ? rest >> 1 : rest;
        if (live && half == 0u && (!kTwoBlocks || (rest & 1u) == 0u))
            mean[ch][run0 + run] = __fmul_rn(r, inv_s);
    }
}

// Out of line: runs once per 256 paths and borrows its registers for a moment (like the ray generator).
static __device__ __noinline__ void fuse_reduce_chunk(const float *slot, unsigned int first, unsigned int valid, unsigned int lane) {
    float *const mean[3] = {c_fuse.mean[0], c_fuse.mean[1], c_fuse.mean[2]};
    const float inv_s = c_fuse.inv_s;
    switch (c_fuse.log2_s) {  // warp-uniform
    case 3: fuse_reduce_chunk_s<3>(slot, mean, first, valid, lane, inv_s); break;
    case 4: fuse_reduce_chunk_s<4>(slot, mean, first, valid, lane, inv_s); break;
    case 5: fuse_reduce_chunk_s<5>(slot, mean, first, valid, lane, inv_s); break;
    case 6: fuse_reduce_chunk_s<6>(slot, mean, first, valid, lane, inv_s); break;
    case 7: fuse_reduce_chunk_s<7>(slot, mean, first, valid, lane, inv_s); break;
    default: fuse_reduce_chunk_s<8>(slot, mean, first, valid, lane, inv_s); break;
    }
}

template <bool GEN, bool FUSE = false> struct PathFeeder {
    const TracePlanes &pl;
    float *ring;
    unsigned long long *counter;
    unsigned int count, lane;
    unsigned int chunk_even, chunk_odd;  // first path of the chunk with even / odd sequence number (>= count: no such chunk)
    unsigned int issued, head;
    unsigned int scratch;                // FUSE: index of this warp's colour scratch in c_fuse.scratch
    bool postponed;                      // FUSE: a chunk claim is waiting for a straggler (warp-uniform)
#ifdef PTB_TEST_POSTPONE
    unsigned int test_postponements = 0u;
#endif

    // GEN: rays are generated straight into the ring (c_gen) and never exist in HBM; pl.ray is unused then.
    __device__ __forceinline__ PathFeeder(const TracePlanes &planes, float *ring_, unsigned long long *counter_, unsigned int count_,
                                          unsigned int lane_)
        : pl(planes), ring(ring_), counter(counter_), count(count_), lane(lane_), issued(0), head(0), scratch(0), postponed(false) {
        chunk_even = chunk_odd = count_;
        if (FUSE)
            scratch = (blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * kFuseScratchFloats;
        for (int b = 0; b < kRingBatches; b++)
            issue(false, 0u);
    }
    // FUSE: where the colour of the path with warp-local sequence number seq is parked (plane c at [c * 2 * kChunkPaths]):
    // one IMAD.WIDE off the constant-bank base, the three stores share the address register
    __device__ __forceinline__ float *park(unsigned int seq) const { return c_fuse.scratch + (scratch + (seq % (2 * kChunkPaths))); }
    __device__ __forceinline__ void reduce_slot(unsigned int parity) {
        const unsigned int first = parity ? chunk_odd : chunk_even;
        if (first < count) {  // warp-uniform
            PTB_CHECK(first % kChunkPaths == 0 && (count - first >= kChunkPaths || ((count - first) & ((1u << c_fuse.log2_s) - 1u)) == 0u));
            __syncwarp();     // the lanes' parked colours are visible to the whole warp
            fuse_reduce_chunk(c_fuse.scratch + (scratch + parity * kChunkPaths), first, count - first < kChunkPaths ? count - first : kChunkPaths, lane);
        }
    }
    // FUSE, after the last path has retired: the (at most two) chunks not yet reduced, oldest first
    __device__ __forceinline__ void finish() {
        const unsigned int chunks = (issued + kChunkBatches - 1) / kChunkBatches;
        if (chunks >= 2u)
            reduce_slot((chunks - 2u) & 1u);
        if (chunks >= 1u)
            reduce_slot((chunks - 1u) & 1u);
    }
    __device__ __forceinline__ unsigned int path_of(unsigned int seq) const {  // warp-local sequence number -> path index
        const unsigned int batch = seq >> 5;
        return (((batch / kChunkBatches) & 1u) ? chunk_odd : chunk_even) + (batch % kChunkBatches) * 32u + (seq & 31u);
    }
    // warp-uniform: every lane commits a (possibly empty) group.  holds / held_seq: this lane holds an unfinished path and its
    // sequence number (FUSE only).  Returns true when a claim that had been postponed has just gone through: lanes that went
    // idle in the meantime ask for a path again.
    __device__ __forceinline__ bool issue(bool holds, unsigned int held_seq) {
        bool rearm = false;
        if (issued % kChunkBatches == 0) {
            if (FUSE && issued >= 2u * kChunkBatches) {
                // The new chunk takes over the scratch slot of the chunk two back, which is reduced first -- once its last path
                // has retired.  A straggler (a path of that chunk still bouncing: only with paths far longer than average)
                // postpones the claim: take() hands out nothing beyond what was issued, lanes that find the ring dry go idle, and
                // the warp tries again at its next swap (there is one: the straggler's own).  In practice the ring's four
                // batches outlast any path (depth 50: ~55 iterations of supply).
                bool straggler = __any_sync(0xffffffffu, holds && held_seq < (issued - kChunkBatches) * 32u);
#ifdef PTB_TEST_POSTPONE  // test build: every other claim is postponed three times although nothing straggles (dry ring, re-arm)
                if (!straggler && ((issued / kChunkBatches) & 1u) == 0u && test_postponements < 3u) {
                    test_postponements++;
                    straggler = true;
                } else if (!straggler) {
                    test_postponements = 0u;
                }
#endif
                if (straggler) {
                    postponed = true;
                    return false;
                }
                rearm = postponed;
                postponed = false;
                reduce_slot((issued / kChunkBatches) & 1u);
            }
            unsigned long long base = 0;
            if (lane == 0)
                base = atomicAdd(counter, static_cast<unsigned long long>(32 * kChunkBatches));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned int first = base < count ? static_cast<unsigned int>(base) : count;
            if ((issued / kChunkBatches) & 1u)
                chunk_odd = first;
            else
                chunk_even = first;
        }
        const unsigned int seq = issued * 32u + lane;
        const unsigned int path = path_of(seq);
        PTB_CHECK(issued * 32u - head <= static_cast<unsigned int>(kRing) - 32u);  // the batch being written is not one still to be read
        if (path < count) {
            const unsigned int s = seq & (kRing - 1);
            if (GEN) {
                generate_ray_to_ring(path, ring + s * kRingEntry);
            } else {
#pragma unroll
                for (int c = 0; c < 6; c++)
                    __pipeline_memcpy_async(ring + c * kRing + s, pl.ray[c] + path, sizeof(float));
            }
        }
        if (!GEN)
            __pipeline_commit();
        issued++;
        return rearm;
    }
    // Called by the whole warp when at least one lane wants a path.  Lanes that get one receive its index and a pointer
    // to its ray in the ring (load_ray(slot, ...)); the caller reads it and then calls refill().
    // seq: the path's warp-local sequence number (FUSE parks its colour by it).  FUSE never hands out a sequence number that has
    // not been issued (the ring can run dry behind a postponed chunk claim).
    __device__ __forceinline__ bool take(bool want, unsigned int wmask, unsigned int &path, unsigned int &seq, const float *&slot) {
        unsigned int lt;
        asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
        seq = head + __popc(wmask & lt);
        path = path_of(seq);
        const bool got = want && path < count && (!FUSE || seq < issued * 32u);
        if (!GEN)
            __pipeline_wait_prior(1);  // everything but the newest batch has landed ...
        __syncwarp();                  // ... and is visible to the other lanes of the warp
        slot = ring + (seq & (kRing - 1)) * (GEN ? kRingEntry : 1);
        PTB_CHECK(!got || (seq < issued * 32u && issued * 32u - seq <= static_cast<unsigned int>(kRing)));  // issued, and not yet overwritten
        head += __popc(wmask);
        if (FUSE)
            head = min(head, issued * 32u);
        return got;
    }
    __device__ __forceinline__ bool refill(bool holds = false, unsigned int held_seq = 0u) {
        __syncwarp();  // ring reads done before a slot can be refilled
        if (issued < head / 32u + kRingBatches)
            return issue(holds, held_seq);
        return false;
    }
};

template <bool GEN> __device__ __forceinline__ void load_ray(const float *slot, float &ox, float &oy, float &oz, float &dx, float &dy, float &dz) {
    if (GEN) {
        const float4 a = *reinterpret_cast<const float4 *>(slot);
        const float2 b = *reinterpret_cast<const float2 *>(slot + 4);
        ox = a.x, oy = a.y, oz = a.z, dx = a.w, dy = b.x, dz = b.y;
    } else {
        ox = slot[0 * kRing], oy = slot[1 * kRing], oz = slot[2 * kRing];
        dx = slot[3 * kRing], dy = slot[4 * kRing], dz = slot[5 * kRing];
    }
}

#ifndef PTB_BLOCKS_PER_SM
#define PTB_BLOCKS_PER_SM 5
#endif
template <int NS, bool EARLY, bool GEN, bool FUSE = false>
__global__ void __launch_bounds__(kTraceThreads, PTB_BLOCKS_PER_SM) trace_paths_kernel(const TracePlanes pl, const float *__restrict__ spheres, unsigned int count,
                                                                       int depth, int nsph, int stride, int light, float scale, float one,
                                                                       unsigned long long *__restrict__ stats, unsigned long long *work_counter) {
    extern __shared__ float4 smem[];
    SceneShared sh;
#ifndef PTB_OLD_LOOP
    // Shared memory: [block-level tables (generic sphere count only)][per warp: ring, then -- 8-sphere kernels -- the warp's own copy
    // of the two lookup tables, so that ring and tables hang off ONE per-warp base register (the block-level table base was
    // rematerialised from S2UR / ULEA every bounce)].
    float *const wsm = reinterpret_cast<float *>(smem + (NS > 0 ? 0 : 2 * nsph)) + (threadIdx.x >> 5) * (kRingFloats + (NS > 0 ? 8 * NS : 0));
    if (NS > 0) {
        sh.center = reinterpret_cast<float4 *>(wsm + kRingFloats);
        sh.color = sh.center + NS;
        const int k = static_cast<int>(threadIdx.x & 31u);
        if (k < NS) {
            sh.center[k] = make_float4(spheres[1 * stride + k], spheres[2 * stride + k], spheres[3 * stride + k], 0.0f);
            sh.color[k] = (EARLY && k == light) ? make_float4(1.0f, 1.0f, 1.0f, 1.0f)
                                                : make_float4(spheres[7 * stride + k], spheres[8 * stride + k], spheres[9 * stride + k], 0.0f);
        }
        __syncwarp();
    } else {
        stage_scene_shared(smem, spheres, nsph, stride, sh, EARLY ? light : -1);
    }
    // Persistent warps with path regeneration.  The loop body is exactly one bounce, so the lanes of a warp may be at
    // different bounces of different paths without diverging.  Lanes whose path finished are ranked by a ballot and take
    // the next consecutive paths of the warp's current chunk, so the rays they fetch and (to within the few paths in
    // flight) the colours they store share cache lines.  See PathFeeder for how paths reach the warp.
    //
    // A lane's whole control state is ONE integer, so that the per-bounce bookkeeping is a compare, a vote and two integer
    // instructions (the first version kept `active` / `want` / `alive` flags in registers and spent ~36 of ~290 instructions per
    // bounce moving them in and out of predicates, profiles/r2_trace_lean_loop.md):
    //   0 <= bounce < depth   the lane holds a path that has done `bounce` bounces
    //   bounce >= depth       the path is finished: depth reached, or settled early (kFin added: the count stays in the low bits)
    //   kNoStore              the lane wants a path and has no colour to store (launch start; FUSE: after a postponed claim)
    //   negative (kIdle + k)  no path, no more work: k <= depth more iterations pass before the warp leaves
    constexpr int kFin = 0x40000000, kNoStore = 0x20000000, kIdle = static_cast<int>(0x80000000u), kCountMask = 0x00ffffff;
    // "throughput is exactly (0,0,0)" is only a stop when the scene allows it (pack_scene_kernel): otherwise compare with NaN
    const unsigned int lane = threadIdx.x & 31u;
    PathFeeder<GEN, FUSE> feed(pl, wsm, work_counter, count, lane);

    PathState p;
    p.ox = p.oy = p.oz = p.dx = p.dy = 0.0f;
    p.dz = 1.0f;
    p.rr = p.rg = p.rb = 1.0f;
    p.alive = true;
    int bounce = kNoStore;
    unsigned int segs = 0;
    unsigned int mine = 0;      // path this lane holds (FUSE: its warp-local sequence number)

    for (;;) {
        const bool want = bounce >= depth;
        const unsigned int wmask = __ballot_sync(0xffffffffu, want);
        if (wmask != 0u) {  // warp-uniform
            if (want && bounce != kNoStore) {  // lanes that finished a path in the previous iteration
                if (FUSE) {                    // parked for the warp's own resolve (fuse_reduce_chunk)
                    float *park = feed.park(mine);
                    PTB_CHECK(feed.path_of(mine) < count && mine < feed.head && mine / kChunkPaths + 2u >= (feed.issued + kChunkBatches - 1) / kChunkBatches);  // its chunk has not been reduced yet: the slot is still its chunk's
                    park[0 * kChunkPaths] = __fmul_rn(p.rr, scale);  // render.cpp:194-196
                    park[2 * kChunkPaths] = __fmul_rn(p.rg, scale);
                    park[4 * kChunkPaths] = __fmul_rn(p.rb, scale);
                } else {
                    PTB_CHECK(mine < count);
                    pl.col[0][mine] = __fmul_rn(p.rr, scale);  // render.cpp:194-196
                    pl.col[1][mine] = __fmul_rn(p.rg, scale);
                    pl.col[2][mine] = __fmul_rn(p.rb, scale);
                }
                segs += static_cast<unsigned int>(bounce & kCountMask);
            }
            unsigned int path, seq;
            const float *slot;
            const bool got = feed.take(want, wmask, path, seq, slot);
            if (got) {
                load_ray<GEN>(slot, p.ox, p.oy, p.oz, p.dx, p.dy, p.dz);
                mine = FUSE ? seq : path;
            }
            const bool rearm = feed.refill(want ? got : bounce >= 0, mine);  // warp-uniform; FUSE only, and next to never
            if (want) {
                bounce = got ? 0 : kIdle;
                p.rr = p.rg = p.rb = 1.0f;
                p.alive = true;
            }
            if (FUSE && rearm) {
                if (bounce < 0)
                    bounce = kNoStore;
                continue;
            }
            if (!__any_sync(0xffffffffu, bounce >= 0))
                break;
        }
        float tmin;
        int idx;
        nearest_hit<NS>(p, nsph, one, kEps, tmin, idx);
        PTB_CHECK(idx >= 0 && idx < nsph);
        if (EARLY) {
            // exact early termination: the light reached (every later factor is exactly 1) or the throughput exactly (+0,+0,+0)
            const bool lit = bounce_and_shade_early(p, tmin, idx, sh, one);
            bool settled = lit;
            if (fmaxf(fmaxf(p.rr, p.rg), p.rb) == c_scene_zero_or_nan)
                settled = true;
            bounce++;
            if (settled)
                bounce |= kFin;  // OR, not +: an idle lane (negative) shades garbage and must stay negative
        } else {
            bounce_and_shade<false>(p, tmin, idx, light, sh);
            bounce++;
        }
    }
    if (FUSE)
        feed.finish();
#else
    stage_scene_shared(smem, spheres, nsph, stride, sh);
    const bool zero_stop = c_scene_zero_stop_ok != 0;

    // Persistent warps with path regeneration.  The loop body is exactly one bounce, so the lanes of a warp may be at
    // different bounces of different paths without diverging.  Lanes whose path finished are ranked by a ballot and take
    // the next consecutive paths of the warp's current chunk, so the rays they fetch and (to within the few paths in
    // flight) the colours they store share cache lines.  See PathFeeder for how paths reach the warp.
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int warp_in_block = threadIdx.x >> 5;
    PathFeeder<GEN, FUSE> feed(pl, reinterpret_cast<float *>(smem + 2 * nsph) + warp_in_block * kRingFloats, work_counter, count, lane);

    PathState p;
    p.ox = p.oy = p.oz = p.dx = p.dy = 0.0f;
    p.dz = 1.0f;
    p.rr = p.rg = p.rb = 1.0f;
    p.alive = true;
    int bounce = 0;
    unsigned int segs = 0;
    unsigned int mine = 0;      // path this lane holds (FUSE: its warp-local sequence number)
    bool active = false;        // this lane holds a real path
    bool want = true;           // this lane needs a (new) path

    for (;;) {
        const unsigned int wmask = __ballot_sync(0xffffffffu, want);
        if (wmask != 0u) {  // warp-uniform
            if (active && want) {  // lanes that finished a path in the previous iteration
                if (FUSE) {        // parked for the warp's own resolve (fuse_reduce_chunk)
                    float *park = feed.park(mine);
                    PTB_CHECK(feed.path_of(mine) < count && mine < feed.head && mine / kChunkPaths + 2u >= (feed.issued + kChunkBatches - 1) / kChunkBatches);  // its chunk has not been reduced yet: the slot is still its chunk's
                    park[0 * kChunkPaths] = __fmul_rn(p.rr, scale);  // render.cpp:194-196
                    park[2 * kChunkPaths] = __fmul_rn(p.rg, scale);
                    park[4 * kChunkPaths] = __fmul_rn(p.rb, scale);
                } else {
                    PTB_CHECK(mine < count);
                    pl.col[0][mine] = __fmul_rn(p.rr, scale);  // render.cpp:194-196
                    pl.col[1][mine] = __fmul_rn(p.rg, scale);
                    pl.col[2][mine] = __fmul_rn(p.rb, scale);
                }
                segs += bounce;
            }
            unsigned int path, seq;
            const float *slot;
            const bool got = feed.take(want, wmask, path, seq, slot);
            if (got) {
                load_ray<GEN>(slot, p.ox, p.oy, p.oz, p.dx, p.dy, p.dz);
                mine = FUSE ? seq : path;
            }
            const bool rearm = feed.refill(want ? got : active, mine);  // warp-uniform; FUSE only, and next to never
            if (want) {
                active = got;
                p.rr = p.rg = p.rb = 1.0f;
                p.alive = true;
                bounce = 0;
                want = false;
            }
            if (FUSE && rearm) {
                want = !active;
                continue;
            }
            if (!__any_sync(0xffffffffu, active))
                break;
        }
        float tmin;
        int idx;
        nearest_hit<NS>(p, nsph, one, kEps, tmin, idx);
        PTB_CHECK(idx >= 0 && idx < nsph);
        bounce_and_shade<EARLY>(p, tmin, idx, light, sh);
        bounce++;
        want = active && ((bounce >= depth) || (EARLY && path_settled(p, zero_stop)));
    }
    if (FUSE)
        feed.finish();
#endif
    if (stats != nullptr) {
        unsigned int w = segs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            w += __shfl_xor_sync(0xffffffffu, w, o);
        if ((threadIdx.x & 31) == 0)
            atomicAdd(stats, static_cast<unsigned long long>(w));
    }
}


// ---- material extension kernel (pt_material.cuh), constant-bank scenes -----------------------------------------------
// Same persistent warps, ring and ballot-ranked regeneration; one iteration = one bounce of material_bounce().
template <int NS, bool GEN, bool FUSE = false>
__global__ void __launch_bounds__(kTraceThreads, 3) trace_materials_kernel(const TracePlanes pl, const float *__restrict__ spheres, unsigned int count,
                                                                           int max_depth, int rr_start, int nsph, int stride, float eps, float one,
                                                                           unsigned long long seed, unsigned long long path0,
                                                                           unsigned long long *__restrict__ stats, unsigned long long *work_counter) {
    extern __shared__ float4 smem[];
    MatShared sh;
    stage_materials_shared(smem, spheres, nsph, stride, sh);
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int warp_in_block = threadIdx.x >> 5;
    PathFeeder<GEN, FUSE> feed(pl, reinterpret_cast<float *>(smem + 3 * nsph) + warp_in_block * kRingFloats, work_counter, count, lane);

    MatPath p;
    p.ox = p.oy = p.oz = p.dx = p.dy = 0.0f;
    p.dz = 1.0f;
    p.tr = p.tg = p.tb = 1.0f;
    p.lr = p.lg = p.lb = 0.0f;
    p.depth = 0;
    unsigned int segs = 0, mine = 0, mine_seq = 0;
    // one integer of lane state, as in trace_paths_kernel: 0 <= state < max_depth running (= p.depth); >= max_depth finished (kFin
    // OR-ed in when the path ended before the cap); kNoStore wants a path, nothing to store; negative idle
    constexpr int kFin = 0x40000000, kNoStore = 0x20000000, kIdle = static_cast<int>(0x80000000u);
    int state = kNoStore;

    for (;;) {
        const bool want = state >= max_depth;
        const unsigned int wmask = __ballot_sync(0xffffffffu, want);
        if (wmask != 0u) {
            if (want && state != kNoStore) {
                if (FUSE) {  // parked for the warp's own resolve (fuse_reduce_chunk)
                    float *park = feed.park(mine_seq);
                    park[0 * kChunkPaths] = p.lr, park[2 * kChunkPaths] = p.lg, park[4 * kChunkPaths] = p.lb;
                } else {
                    PTB_CHECK(mine < count);
                    pl.col[0][mine] = p.lr;
                    pl.col[1][mine] = p.lg;
                    pl.col[2][mine] = p.lb;
                }
            }
            unsigned int path, seq;
            const float *slot;
            const bool got = feed.take(want, wmask, path, seq, slot);
            if (got) {
                load_ray<GEN>(slot, p.ox, p.oy, p.oz, p.dx, p.dy, p.dz);
                mine = path;
                mine_seq = seq;
            }
            const bool rearm = feed.refill(want ? got : state >= 0, mine_seq);  // warp-uniform; FUSE only, and next to never
            if (want) {
                state = got ? 0 : kIdle;
                p.tr = p.tg = p.tb = 1.0f;
                p.lr = p.lg = p.lb = 0.0f;
            }
            if (FUSE && rearm) {
                if (state < 0)
                    state = kNoStore;
                continue;
            }
            if (!__any_sync(0xffffffffu, state >= 0))
                break;
        }
        if (state >= 0) {  // idle lanes of a finished range wait; live ones diverge by material inside
            segs++;
            p.depth = state;
            // RNG key: the global path index (a strided launch walks a dense frame of every k-th column, pt_raygen.cuh)
            const unsigned long long pid = (GEN && c_gen.x_step > 1) ? strided_global_path(c_gen, mine) : path0 + mine;
            const bool ended = material_bounce<NS>(p, nsph, one, eps, rr_start, seed, pid, sh);
            state = ended ? (p.depth | kFin) : p.depth;
        }
    }
    if (FUSE)
        feed.finish();
    if (stats != nullptr) {
        unsigned int w = segs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0)
            atomicAdd(stats, static_cast<unsigned long long>(w));
    }
}

// ---- material extension kernel, BVH scenes: warp-local wavefront -------------------------------------------------------
// A tree walk takes 28 steps on average in the C4 scene but ten times that for the unluckiest ray of a warp, so a warp
// that walks 32 rays in lock step and shades them together keeps 3 of its 32 lanes busy (ncu, profiles/r2_c4_v1_*).
// Here lanes are decoupled from paths.  Each warp owns a pool of kPool path slots in shared memory and two queues of slot
// numbers: rays waiting for a tree walk (tq) and hits waiting for shading (sq).  A lane takes any waiting ray, walks the
// tree, drops the hit into sq and takes the next ray; when 32 hits have gathered the whole warp shades them at once
// (materials, Russian roulette, regeneration of ended paths, the brute-force pass over the huge spheres) and the
// continuing rays go back to tq.  With kPool = 64 every lane always finds a ray while work lasts: 32 rays in flight, the
// other 32 slots split between the two queues, and shading runs exactly when tq is empty.
// Queue traffic is batched: the walk loop runs until kRefillLanes lanes have finished before the warp stops to requeue.
constexpr int kPool = 64;
#ifndef PTB_SHORT_STACK
#define PTB_SHORT_STACK 12
#endif
constexpr int kShortStack = PTB_SHORT_STACK;  // stack entries per lane in shared memory; deeper ones overflow to local memory
#ifndef PTB_REFILL_LANES
#define PTB_REFILL_LANES 8
#endif
constexpr int kRefillLanes = PTB_REFILL_LANES;
#ifndef PTB_LEAF_BATCH
#define PTB_LEAF_BATCH 4
#endif
constexpr int kLeafBatch = PTB_LEAF_BATCH;

struct WarpPool {
    float ray[6][kPool];  // ox oy oz dx dy dz
    float thr[3][kPool];  // throughput
    float rad[3][kPool];  // radiance gathered so far
    int depth[kPool];     // < 0: slot holds no path yet
    unsigned int path[kPool];
    float tmin[kPool];    // best hit so far (brute-force list, then the tree)
    int idx[kPool];
    unsigned char tq[kPool], sq[kPool];
};

// Traversal stack of one lane: kShortStack entries in shared memory (entry k of lane t at s32 + 4 * (k * kTraceThreads + t):
// conflict-free columns), deeper entries in local memory (deep, unbalanced trees only: a branch, not predicated code, because it
// next to never runs).  A traversal step either pushes (both children hit), pops (none hit) or neither, so push and pop share one
// address; the base is a 32-bit shared-window address made opaque to the compiler, which otherwise rematerialises it from
// S2R / S2UR / ULEA at every step (6 instructions).
struct HybridStack {
    unsigned int s32;  // shared-window byte address of this lane's column
    int *ovf;          // local memory, entries kShortStack and up
    int sp;
    __device__ __forceinline__ void init(int *column, int *overflow) {
        s32 = static_cast<unsigned int>(__cvta_generic_to_shared(column));
        asm volatile("mov.u32 %0, %0;" : "+r"(s32));
        ovf = overflow;
        sp = 0;
    }
    // push x (push), or pop into the return value (pop; -1 when the stack is empty), or return `otherwise`
    __device__ __forceinline__ int step(bool push, bool pop, int x, int otherwise) {
        const int k = pop ? sp - 1 : sp;  // the entry touched
        int v = otherwise;
        if (pop)
            v = -1;
        const unsigned int a = s32 + static_cast<unsigned int>(k) * (kTraceThreads * 4u);
        if (k < kShortStack) {
            if (push)
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
            if (pop && k >= 0)
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
        } else {  // rare
            PTB_CHECK(k < kBvhStack);
            if (push)
                ovf[k - kShortStack] = x;
            if (pop)
                v = ovf[k - kShortStack];
        }
        sp = push ? sp + 1 : ((pop && k >= 0) ? k : sp);
        return v;
    }
};

struct PoolRay {
    const WarpPool *pool;
    int slot;
    __device__ __forceinline__ float ox() const { return pool->ray[0][slot]; }
    __device__ __forceinline__ float oy() const { return pool->ray[1][slot]; }
    __device__ __forceinline__ float oz() const { return pool->ray[2][slot]; }
    __device__ __forceinline__ float dx() const { return pool->ray[3][slot]; }
    __device__ __forceinline__ float dy() const { return pool->ray[4][slot]; }
    __device__ __forceinline__ float dz() const { return pool->ray[5][slot]; }
};

static __device__ __noinline__ void generate_ray_to_pool(unsigned int path, float *slot) {
    float r[6];
    generate_ray(c_gen, static_cast<long long>(path), r);
#pragma unroll
    for (int c = 0; c < 6; c++)
        slot[c * kPool] = r[c];
}

#ifndef PTB_BVH_MAX_CHUNK
#define PTB_BVH_MAX_CHUNK 64
#endif
#ifndef PTB_BVH_BLOCKS_PER_SM
#define PTB_BVH_BLOCKS_PER_SM 4
#endif
template <bool GEN>
__global__ void __launch_bounds__(kTraceThreads, PTB_BVH_BLOCKS_PER_SM)
    trace_materials_bvh_kernel(const TracePlanes pl, unsigned int count, int max_depth, int rr_start, int nbig, float eps, float one,
                               unsigned long long seed, unsigned long long path0, unsigned long long *__restrict__ stats, const BvhScene bvh,
                               unsigned long long *work_counter, unsigned int chunk) {
    extern __shared__ float4 smem[];
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int warp_in_block = threadIdx.x >> 5;
    unsigned int lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    WarpPool &pool = reinterpret_cast<WarpPool *>(smem)[warp_in_block];
    const MatShared nosh = {nullptr, nullptr, nullptr};

    // every slot starts in the shading queue as "no path yet": the first two shading rounds just fetch paths
    pool.depth[lane] = -1, pool.depth[lane + 32] = -1;
    pool.sq[lane] = static_cast<unsigned char>(lane), pool.sq[lane + 32] = static_cast<unsigned char>(lane + 32);
    int sq_count = kPool, tq_count = 0;
    unsigned int next = 0, chunk_end = 0;  // the warp's current chunk of paths: [next, chunk_end)
    __syncwarp();

    // node >= 0: walking, at that internal node; kWalked: holds a ray whose walk is over (hit not yet deposited); kNoRay: free lane
    constexpr int kWalked = -1, kNoRay = -2;
    int slot = 0, node = kNoRay, idx = 0;
    int pend = -1;     // a hit leaf whose exact test waits for company (kLeafBatch lanes) or for the next requeue stop
    float tmin = kMiss;
    BvhRay r = {};
    int stack_overflow[kBvhStack - kShortStack];
    HybridStack st;
    st.init(reinterpret_cast<int *>(reinterpret_cast<WarpPool *>(smem) + kWarpsPerBlock) + threadIdx.x, stack_overflow);
    unsigned int segs = 0;

    for (;;) {
        // 1. leaf tests still pending, then lanes whose walk ended deposit their hit
        if (pend >= 0) {
            bvh_leaf(bvh, pend, pool.ray[0][slot], pool.ray[1][slot], pool.ray[2][slot], pool.ray[3][slot], pool.ray[4][slot], pool.ray[5][slot], eps,
                     tmin, idx);
            pend = -1;
        }
        const bool fin = node == kWalked;
        const unsigned int fmask = __ballot_sync(0xffffffffu, fin);
        if (fmask != 0u) {
            if (fin) {
                PTB_CHECK(slot >= 0 && slot < kPool && sq_count + __popc(fmask & lt) < kPool);
                pool.tmin[slot] = tmin;
                pool.idx[slot] = idx;
                pool.sq[sq_count + __popc(fmask & lt)] = static_cast<unsigned char>(slot);
                node = kNoRay;
            }
            sq_count += __popc(fmask);
            PTB_CHECK(sq_count + tq_count + __popc(__ballot_sync(0xffffffffu, node != kNoRay)) <= kPool);  // a slot is in at most one place (fewer at the very end)
            __syncwarp();
        }
        const unsigned int busy = __ballot_sync(0xffffffffu, node != kNoRay);
        // 2. a full warp of hits (or the last few): shade, regenerate, brute-force pass, back to tq
        if (sq_count >= 32 || (sq_count > 0 && busy == 0u && tq_count == 0)) {
            const int n_take = sq_count < 32 ? sq_count : 32;
            const bool has = static_cast<int>(lane) < n_take;
            int s2 = 0;
            if (has)
                s2 = pool.sq[sq_count - n_take + lane];
            PTB_CHECK(s2 >= 0 && s2 < kPool);
            sq_count -= n_take;
            MatPath p;
            p.ox = p.oy = p.oz = p.dx = p.dy = 0.0f;
            p.dz = 1.0f;
            p.tr = p.tg = p.tb = 1.0f;
            p.lr = p.lg = p.lb = 0.0f;
            p.depth = 0;
            unsigned int mypath = 0;
            bool ended = true;
            if (has && pool.depth[s2] >= 0) {
                p.ox = pool.ray[0][s2], p.oy = pool.ray[1][s2], p.oz = pool.ray[2][s2];
                p.dx = pool.ray[3][s2], p.dy = pool.ray[4][s2], p.dz = pool.ray[5][s2];
                p.tr = pool.thr[0][s2], p.tg = pool.thr[1][s2], p.tb = pool.thr[2][s2];
                p.lr = pool.rad[0][s2], p.lg = pool.rad[1][s2], p.lb = pool.rad[2][s2];
                p.depth = pool.depth[s2];
                mypath = pool.path[s2];
                segs++;
                const unsigned long long pid = (GEN && c_gen.x_step > 1) ? strided_global_path(c_gen, mypath) : path0 + mypath;
                ended = material_shade<true>(p, pool.tmin[s2], pool.idx[s2], rr_start, seed, pid, nosh, bvh) || p.depth >= max_depth;
                if (ended) {
                    PTB_CHECK(mypath < count);
                    pl.col[0][mypath] = p.lr;
                    pl.col[1][mypath] = p.lg;
                    pl.col[2][mypath] = p.lb;
                }
            }
            // ended paths are replaced by the next consecutive paths of the warp's chunk (ballot-ranked)
            const bool need = has && ended;
            const unsigned int nmask = __ballot_sync(0xffffffffu, need);
            bool got = false;
            if (nmask != 0u) {
                const unsigned int n_need = __popc(nmask), rank = __popc(nmask & lt);
                const unsigned int avail = chunk_end - next;
                unsigned int fresh = count, fresh_end = count;
                if (n_need > avail) {  // warp-uniform: claim the next chunk
                    unsigned long long base = 0;
                    if (lane == 0)
                        base = atomicAdd(work_counter, static_cast<unsigned long long>(chunk));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    fresh = base < count ? static_cast<unsigned int>(base) : count;
                    fresh_end = (count - fresh < chunk) ? count : fresh + chunk;
                }
                const unsigned int path = rank < avail ? next + rank : fresh + (rank - avail);
                got = need && (rank < avail || path < fresh_end);
                PTB_CHECK(!got || path < count);
                if (n_need > avail) {
                    next = fresh + (n_need - avail);
                    chunk_end = fresh_end;
                    if (next > chunk_end)
                        next = chunk_end;
                } else {
                    next += n_need;
                }
                if (got) {
                    if (GEN) {
                        generate_ray_to_pool(path, &pool.ray[0][s2]);
                        p.ox = pool.ray[0][s2], p.oy = pool.ray[1][s2], p.oz = pool.ray[2][s2];
                        p.dx = pool.ray[3][s2], p.dy = pool.ray[4][s2], p.dz = pool.ray[5][s2];
                    } else {
                        p.ox = pl.ray[0][path], p.oy = pl.ray[1][path], p.oz = pl.ray[2][path];
                        p.dx = pl.ray[3][path], p.dy = pl.ray[4][path], p.dz = pl.ray[5][path];
                    }
                    p.tr = p.tg = p.tb = 1.0f;
                    p.lr = p.lg = p.lb = 0.0f;
                    p.depth = 0;
                    mypath = path;
                }
            }
            const bool cont = has && (!ended || got);
            if (cont) {  // all lanes except at the very end of the launch
                float t = kMiss;
                int i = 0;
                if (nbig > 0) {
                    PathState ray;
                    ray.ox = p.ox, ray.oy = p.oy, ray.oz = p.oz, ray.dx = p.dx, ray.dy = p.dy, ray.dz = p.dz;
                    nearest_hit<0>(ray, nbig, one, eps, t, i);
                    i = (t < kMiss) ? __ldg(bvh.big_index + i) : 0;
                }
                pool.ray[0][s2] = p.ox, pool.ray[1][s2] = p.oy, pool.ray[2][s2] = p.oz;
                pool.ray[3][s2] = p.dx, pool.ray[4][s2] = p.dy, pool.ray[5][s2] = p.dz;
                pool.thr[0][s2] = p.tr, pool.thr[1][s2] = p.tg, pool.thr[2][s2] = p.tb;
                pool.rad[0][s2] = p.lr, pool.rad[1][s2] = p.lg, pool.rad[2][s2] = p.lb;
                pool.depth[s2] = p.depth;
                pool.path[s2] = mypath;
                pool.tmin[s2] = t;
                pool.idx[s2] = i;
            }
            const unsigned int cmask = __ballot_sync(0xffffffffu, cont);
            PTB_CHECK(tq_count + __popc(cmask) <= kPool);
            if (cont)
                pool.tq[tq_count + __popc(cmask & lt)] = static_cast<unsigned char>(s2);
            tq_count += __popc(cmask);
            __syncwarp();
        }
        // 3. idle lanes take waiting rays
        if (tq_count > 0 && busy != 0xffffffffu) {
            const unsigned int imask = ~busy;
            const int rank = __popc(imask & lt);
            float ox = 0.0f, oy = 0.0f, oz = 0.0f, dx = 0.0f, dy = 0.0f, dz = 1.0f;
            bool exact_loop = false;  // a ray the quantised walk does not cover (pt_bvh.cuh): every sphere, by the whole warp
            if (node == kNoRay && rank < tq_count) {
                slot = pool.tq[tq_count - 1 - rank];
                PTB_CHECK(slot >= 0 && slot < kPool);
                ox = pool.ray[0][slot], oy = pool.ray[1][slot], oz = pool.ray[2][slot];
                dx = pool.ray[3][slot], dy = pool.ray[4][slot], dz = pool.ray[5][slot];
                tmin = pool.tmin[slot];
                idx = pool.idx[slot];
                node = kWalked;
                st.sp = 0;
                if (bvh.n_small == 1) {
                    bvh_leaf(bvh, ~bvh.only_leaf, ox, oy, oz, dx, dy, dz, eps, tmin, idx);
                } else if (bvh.n_small > 1) {
                    r = bvh_ray(bvh, ox, oy, oz, dx, dy, dz, tmin);
                    exact_loop = r.far_origin;
                    if (!exact_loop)
                        node = 0;
                }
            }
            for (unsigned int todo = __ballot_sync(0xffffffffu, exact_loop); todo != 0u; todo &= todo - 1u) {
                const int leader = __ffs(todo) - 1;
                float t = __shfl_sync(0xffffffffu, tmin, leader);
                int i = __shfl_sync(0xffffffffu, idx, leader);
                bvh_all_leaves_warp(bvh, lane, __shfl_sync(0xffffffffu, ox, leader), __shfl_sync(0xffffffffu, oy, leader),
                                    __shfl_sync(0xffffffffu, oz, leader), __shfl_sync(0xffffffffu, dx, leader),
                                    __shfl_sync(0xffffffffu, dy, leader), __shfl_sync(0xffffffffu, dz, leader), eps, t, i);
                if (static_cast<int>(lane) == leader) {
                    tmin = t;
                    idx = i;
                }
            }
            const int taken = __popc(imask);
            tq_count -= taken < tq_count ? taken : tq_count;
            __syncwarp();
        }
        // 4. walk until kRefillLanes more lanes have finished (or nobody walks any more)
        const unsigned int act = __ballot_sync(0xffffffffu, node >= 0);
        if (act == 0u) {
            if (__ballot_sync(0xffffffffu, node != kNoRay) == 0u && sq_count == 0 && tq_count == 0)
                break;
            continue;
        }
        const int stop_at = __popc(act) > kRefillLanes ? __popc(act) - kRefillLanes : 0;
        const PoolRay pray = {&pool, slot};
        do {
            int leaf_a = -1, leaf_b = -1;
            if (node >= 0)
                bvh_step_boxes(bvh, r, tmin, node, leaf_a, leaf_b, st);
            // Exact leaf tests are rare per lane (0.7 per ray) but a warp of 28 walkers meets one every other step, and run
            // at once each costs the whole warp ~30 instructions and a memory round trip for one lane's benefit.  So a hit
            // leaf waits in `pend` until kLeafBatch lanes have one (or the walk loop stops); the lane walks on with its
            // old tmin meanwhile, which only culls less.  A lane that already holds one tests the older at once.
            if (leaf_b >= 0)
                bvh_leaf(bvh, leaf_b, pray.ox(), pray.oy(), pray.oz(), pray.dx(), pray.dy(), pray.dz(), eps, tmin, idx);
            if (leaf_a >= 0 && pend >= 0)
                bvh_leaf(bvh, pend, pray.ox(), pray.oy(), pray.oz(), pray.dx(), pray.dy(), pray.dz(), eps, tmin, idx);
            pend = leaf_a >= 0 ? leaf_a : pend;
            if (__popc(__ballot_sync(0xffffffffu, pend >= 0)) >= kLeafBatch) {
                if (pend >= 0)
                    bvh_leaf(bvh, pend, pray.ox(), pray.oy(), pray.oz(), pray.dx(), pray.dy(), pray.dz(), eps, tmin, idx);
                pend = -1;
            }
        } while (__popc(__ballot_sync(0xffffffffu, node >= 0)) > stop_at);
    }
    if (stats != nullptr) {
        unsigned int w = segs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0)
            atomicAdd(stats, static_cast<unsigned long long>(w));
    }
}

// ---- host side --------------------------------------------------------------------------------------
namespace {

struct DeviceState {
    bool init = false;
    int device = 0;
    int sm_count = 0;
    int blocks_per_sm[16] = {};  // [FUSE?][GEN?][NS8?][EARLY?]
    SceneConst *scene_alias = nullptr;
    int *zero_ok_alias = nullptr;
    float *zero_or_nan_alias = nullptr;
    unsigned long long *work_counter = nullptr;  // chunk dispenser of the persistent kernels (reset before every launch)
    float *fuse_scratch = nullptr;     // fused resolve: kFuseScratchFloats per warp of the largest grid launched so far
    size_t fuse_scratch_warps = 0;
    cudaEvent_t scene_free = nullptr;  // recorded after the last kernel that reads the staged scene
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
};

constexpr int kMaxDevices = 64;
DeviceState g_dev[kMaxDevices];
std::mutex g_mu;

template <int NS, bool EARLY, bool GEN, bool FUSE> cudaError_t occupancy(int *out, size_t smem) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, trace_paths_kernel<NS, EARLY, GEN, FUSE>, kTraceThreads, smem);
}

// Fused resolve: points c_fuse at the run means of the launch that starts at path `a` of the call's range and makes sure the
// per-warp colour scratch covers `grid` blocks.  Caller holds g_mu and has ordered the stream behind earlier launches.
cudaError_t stage_fuse(DeviceState &s, cudaStream_t stream, const FuseTarget &f, int64_t a, int samples, int grid) {
    const size_t warps = static_cast<size_t>(grid) * kWarpsPerBlock;
    if (warps > s.fuse_scratch_warps) {
        cudaError_t e = cudaSuccess;
        if (s.fuse_scratch != nullptr && (e = cudaFree(s.fuse_scratch)) != cudaSuccess)  // waits for the kernels that use it
            return e;
        s.fuse_scratch = nullptr, s.fuse_scratch_warps = 0;
        if ((e = cudaMalloc(reinterpret_cast<void **>(&s.fuse_scratch), warps * kFuseScratchFloats * sizeof(float))) != cudaSuccess)
            return e;
        s.fuse_scratch_warps = warps;
    }
    unsigned int lg = 0;
    while ((1 << lg) < samples)
        lg++;
    FuseOut o;
    for (int c = 0; c < 3; c++)
        o.mean[c] = f.means + c * f.n_runs + a / samples;
    o.scratch = s.fuse_scratch;
    o.log2_s = lg;
    o.inv_s = 1.0f / static_cast<float>(samples);
    return cudaMemcpyToSymbolAsync(c_fuse, &o, sizeof o, 0, cudaMemcpyHostToDevice, stream);
}

// Stages the ray generator's parameters (fused generate-and-trace launches); caller holds g_mu and has ordered the stream.
cudaError_t stage_gen(cudaStream_t stream, const RayGenSource &g) {
    return cudaMemcpyToSymbolAsync(c_gen, &g, sizeof g, 0, cudaMemcpyHostToDevice, stream);
}

cudaError_t ensure_device_state(DeviceState **out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev < 0 || dev >= kMaxDevices)
        return cudaErrorInvalidDevice;
    DeviceState &s = g_dev[dev];
    if (!s.init) {
        if ((e = cudaDeviceGetAttribute(&s.sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
            return e;
        if ((e = cudaGetSymbolAddress(reinterpret_cast<void **>(&s.scene_alias), c_scene)) != cudaSuccess)
            return e;
        if ((e = cudaGetSymbolAddress(reinterpret_cast<void **>(&s.zero_ok_alias), c_scene_zero_stop_ok)) != cudaSuccess)
            return e;
        if ((e = cudaGetSymbolAddress(reinterpret_cast<void **>(&s.zero_or_nan_alias), c_scene_zero_or_nan)) != cudaSuccess)
            return e;
        if ((e = cudaEventCreateWithFlags(&s.scene_free, cudaEventDisableTiming)) != cudaSuccess)
            return e;
        if ((e = cudaMalloc(reinterpret_cast<void **>(&s.work_counter), sizeof(unsigned long long))) != cudaSuccess)
            return e;
        s.device = dev;
        s.init = true;
    }
    *out = &s;
    return cudaSuccess;
}

// The staged scene, the generator block and the chunk dispenser belong to the CURRENT device: a stream of another device
// would run the kernels there against this device's staging.  Refused instead of silently misbehaving.
cudaError_t check_stream_device(const DeviceState &s, cudaStream_t stream) {
    if (stream == nullptr || stream == cudaStreamLegacy || stream == cudaStreamPerThread)
        return cudaSuccess;
    int sd = -1;
    if (cudaStreamGetDevice(stream, &sd) != cudaSuccess) {
        cudaGetLastError();
        return cudaErrorInvalidResourceHandle;
    }
    return sd == s.device ? cudaSuccess : cudaErrorInvalidDevice;
}

template <int NS, bool EARLY, bool GEN, bool FUSE = false>
cudaError_t launch_trace(DeviceState &s, cudaStream_t stream, const float *rays, const float *spheres, float *colors, int64_t n,
                         int64_t first, int64_t count, const PtParams &p, unsigned long long *stats, const RayGenSource *gen,
                         const FuseTarget *fuse = nullptr) {
#ifndef PTB_OLD_LOOP
    // 8-sphere kernels: every warp holds its own copy of the lookup tables behind its ring (trace_paths_kernel)
    const size_t smem = NS > 0 ? sizeof(float) * (kRingFloats + 8 * NS) * kWarpsPerBlock
                               : sizeof(float4) * 2 * static_cast<size_t>(p.sphere_count) + sizeof(float) * kRingFloats * kWarpsPerBlock;
#else
    const size_t smem = sizeof(float4) * 2 * static_cast<size_t>(p.sphere_count) + sizeof(float) * kRingFloats * kWarpsPerBlock;
#endif
    int &occ = s.blocks_per_sm[(FUSE ? 8 : 0) + (GEN ? 4 : 0) + (NS > 0 ? 2 : 0) + (EARLY ? 1 : 0)];
    if (occ == 0 || NS == 0) {
        cudaError_t e = cudaSuccess;
        if (smem > 48 * 1024 &&  // 769..1024 spheres: opt in to more than the default 48 KB of dynamic shared memory
            (e = cudaFuncSetAttribute(trace_paths_kernel<NS, EARLY, GEN, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess)
            return e;
        if ((e = occupancy<NS, EARLY, GEN, FUSE>(&occ, smem)) != cudaSuccess)
            return e;
        if (occ < 1) {  // the kernel cannot be resident at all with this much shared memory: fail loudly, do not guess a grid
            occ = 0;
            return cudaErrorLaunchOutOfResources;
        }
    }
    const int64_t cap = static_cast<int64_t>(s.sm_count) * occ;
    constexpr int64_t kMaxPerLaunch = 1LL << 30;  // 32-bit path indices inside the kernel
    for (int64_t a = first; a < first + count; a += kMaxPerLaunch) {
        const int64_t m = (first + count - a < kMaxPerLaunch) ? first + count - a : kMaxPerLaunch;
        TracePlanes pl;
        for (int c = 0; c < 6; c++)
            pl.ray[c] = GEN ? nullptr : rays + c * n + a;
        for (int c = 0; c < 3; c++)
            pl.col[c] = FUSE ? nullptr : colors + c * n + a;
        const int64_t need = (m + kTraceThreads - 1) / kTraceThreads;
        const int grid = static_cast<int>(need < cap ? need : cap);
        cudaError_t e = cudaMemsetAsync(s.work_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess)
            return e;
        if (GEN) {  // element i of this launch is element (a - first) + i of the generator's range
            RayGenSource g = *gen;
            const int64_t off = a - first;
            g = make_raygen_source_shifted(g, off, m);
            if ((e = stage_gen(stream, g)) != cudaSuccess)
                return e;
        }
        if (FUSE && (e = stage_fuse(s, stream, *fuse, a - first, p.samples, grid)) != cudaSuccess)
            return e;
        trace_paths_kernel<NS, EARLY, GEN, FUSE><<<grid, kTraceThreads, smem, stream>>>(pl, spheres, static_cast<unsigned int>(m), p.depth,
                                                                                  p.sphere_count, p.sphere_stride, p.light_index, p.emission_scale,
                                                                                  1.0f, stats, s.work_counter);
        e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

}  // namespace

bool fuse_supported(int samples) { return samples >= 8 && samples <= kChunkPaths && (samples & (samples - 1)) == 0; }

cudaError_t trace_paths(cudaStream_t stream, const PtParams &p, const float *rays, const float *spheres, float *colors, int64_t n,
                        int64_t first, int64_t count, unsigned long long *stats, const RayGenSource *gen, const FuseTarget *fuse) {
    if (count <= 0)
        return cudaSuccess;
    if (fuse != nullptr && (gen == nullptr || !fuse_supported(p.samples) || first % p.samples != 0 || count % p.samples != 0))
        return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceState *s = nullptr;
    cudaError_t e = ensure_device_state(&s);
    if (e != cudaSuccess || (e = check_stream_device(*s, stream)) != cudaSuccess)
        return e;
    // The constant-bank scene is a per-device singleton: a launch sequence on another stream must
    // wait until the previous sequence has finished reading it.
    if (s->have_last && s->last_stream != stream) {
        if ((e = cudaStreamWaitEvent(stream, s->scene_free, 0)) != cudaSuccess)
            return e;
    }
    pack_scene_kernel<<<1, 128, 0, stream>>>(spheres, p.sphere_count, p.sphere_stride, s->scene_alias, s->zero_ok_alias, s->zero_or_nan_alias);
    if ((e = cudaGetLastError()) != cudaSuccess)
        return e;
    // Early termination and the fixed-depth loop give identical bits, so which one runs is a pure performance choice:
    // lock-step lanes swap paths once per `depth` iterations (119-123 Gsegments/s), regenerating lanes every iteration
    // (103-109 Gsegments/s) but they skip the settled part of every path: 21 % of the segments at depth 5, 13 % at depth 4.
    // Measured on B200 (tools/early_crossover.py, 132.7 M paths): depth 3: 3.72 vs 3.50 ms (lock step wins), depth 4: 4.52 vs
    // 4.54 (a tie), depth 5: 5.08 vs 5.58, depth 6: 5.56 vs 6.62, depth 50: 18.6 vs 100.8 -> regeneration from depth 5 on.
    static const int min_early_depth = [] {  // PTB200_EARLY_FROM_DEPTH overrides the crossover (experiments); values < 1 are ignored
        const char *e = getenv("PTB200_EARLY_FROM_DEPTH");
        const int v = e ? atoi(e) : 0;
        return v >= 1 ? v : 5;
    }();
    const bool early = !(p.flags & PTB200_F_FIXED_DEPTH) && p.depth >= min_early_depth;
    if (fuse != nullptr)
        e = p.sphere_count == 8 ? (early ? launch_trace<8, true, true, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen, fuse)
                                         : launch_trace<8, false, true, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen, fuse))
                                : (early ? launch_trace<0, true, true, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen, fuse)
                                         : launch_trace<0, false, true, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen, fuse));
    else if (p.sphere_count == 8)
        e = early ? (gen ? launch_trace<8, true, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen)
                         : launch_trace<8, true, false>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen))
                  : (gen ? launch_trace<8, false, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen)
                         : launch_trace<8, false, false>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen));
    else
        e = early ? (gen ? launch_trace<0, true, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen)
                         : launch_trace<0, true, false>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen))
                  : (gen ? launch_trace<0, false, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen)
                         : launch_trace<0, false, false>(*s, stream, rays, spheres, colors, n, first, count, p, stats, gen));
    if (e != cudaSuccess)
        return e;
    if ((e = cudaEventRecord(s->scene_free, stream)) != cudaSuccess)
        return e;
    s->last_stream = stream;
    s->have_last = true;
    return cudaSuccess;
}


cudaError_t trace_materials(cudaStream_t stream, const PtParams &p_in, const PtMaterialParams &mp, const float *rays, const float *spheres_in,
                            float *colors, int64_t n, int64_t first, int64_t count, uint64_t path0, unsigned long long *stats, const PtBvh *tree,
                            const RayGenSource *gen, const FuseTarget *fuse) {
    if (count <= 0)
        return cudaSuccess;
    if (fuse != nullptr && (gen == nullptr || tree != nullptr || !fuse_supported(p_in.samples) || first % p_in.samples != 0 || count % p_in.samples != 0))
        return cudaErrorInvalidValue;
    // With a tree, the constant bank and the kernel's brute-force loop see only the huge spheres (compacted SoA, stride 1024).
    PtParams p = p_in;
    const float *spheres = spheres_in;
    BvhScene bvh = {};
    if (tree != nullptr) {
        bvh = bvh_scene(tree);
        spheres = bvh_big_soa(tree);
        p.sphere_count = bvh_big_count(tree);
        p.sphere_stride = 1024;
    }
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceState *s = nullptr;
    cudaError_t e = ensure_device_state(&s);
    if (e != cudaSuccess || (e = check_stream_device(*s, stream)) != cudaSuccess)
        return e;
    if (s->have_last && s->last_stream != stream) {
        if ((e = cudaStreamWaitEvent(stream, s->scene_free, 0)) != cudaSuccess)
            return e;
    }
    if (p.sphere_count > 0) {
        pack_scene_kernel<<<1, 128, 0, stream>>>(spheres, p.sphere_count, p.sphere_stride, s->scene_alias, s->zero_ok_alias, s->zero_or_nan_alias);
        if ((e = cudaGetLastError()) != cudaSuccess)
            return e;
    }
    {   // the material RNG's round keys (same stream ordering as the scene: the previous launch sequence has been waited for)
        const PhiloxKeys keys = philox_keys(mp.seed);
        if ((e = cudaMemcpyToSymbolAsync(c_mat_keys, &keys, sizeof keys, 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess)
            return e;
    }
    const bool use_tree = tree != nullptr;
    const size_t smem = use_tree ? sizeof(WarpPool) * kWarpsPerBlock + sizeof(int) * kShortStack * kTraceThreads
                                 : sizeof(float4) * 3 * static_cast<size_t>(p.sphere_count) + sizeof(float) * kRingFloats * kWarpsPerBlock;
    const bool ten = !use_tree && (p.sphere_count == 9 || p.sphere_count == 10);  // smallpt's scene: unrolled pairs (index 9 is padding)
    if (use_tree && smem > 48 * 1024) {  // only with non-default pool / stack sizes: opt in to more than 48 KB per block
        if ((e = cudaFuncSetAttribute(trace_materials_bvh_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess ||
            (e = cudaFuncSetAttribute(trace_materials_bvh_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess)
            return e;
    }
    if (!use_tree && !ten && smem > 48 * 1024) {  // 513..1024 spheres: opt in to more than the default 48 KB of dynamic shared memory
        if ((e = cudaFuncSetAttribute(trace_materials_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess ||
            (e = cudaFuncSetAttribute(trace_materials_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess)
            return e;
    }
    int occ = 0;
    // the fused-generation variants have the same resource footprint as the SoA ones (the generator is an out-of-line call)
    if ((e = use_tree ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_materials_bvh_kernel<false>, kTraceThreads, smem)
              : ten   ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_materials_kernel<10, false>, kTraceThreads, smem)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_materials_kernel<0, false>, kTraceThreads, smem)) != cudaSuccess)
        return e;
    if (occ < 1)  // cannot be resident at all: fail loudly instead of guessing a grid
        return cudaErrorLaunchOutOfResources;
    const int64_t cap = static_cast<int64_t>(s->sm_count) * occ;
    constexpr int64_t kMaxPerLaunch = 1LL << 30;
    for (int64_t a = first; a < first + count; a += kMaxPerLaunch) {
        const int64_t m = (first + count - a < kMaxPerLaunch) ? first + count - a : kMaxPerLaunch;
        TracePlanes pl;
        for (int c = 0; c < 6; c++)
            pl.ray[c] = rays + c * n + a;
        for (int c = 0; c < 3; c++)
            pl.col[c] = fuse ? nullptr : colors + c * n + a;
        // the wavefront kernel keeps kPool paths per warp in flight, the lock-step kernels one per lane
        const int64_t per_block = use_tree ? static_cast<int64_t>(kPool) * kWarpsPerBlock : kTraceThreads;
        const int64_t need = (m + per_block - 1) / per_block;
        const int grid = static_cast<int>(need < cap ? need : cap);
        if ((e = cudaMemsetAsync(s->work_counter, 0, sizeof(unsigned long long), stream)) != cudaSuccess)
            return e;
        const unsigned int mm = static_cast<unsigned int>(m);
        const unsigned long long pp = path0 + static_cast<uint64_t>(a - first);
        if (gen != nullptr) {
            pl = TracePlanes{{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, {pl.col[0], pl.col[1], pl.col[2]}};
            if ((e = stage_gen(stream, make_raygen_source_shifted(*gen, a - first, m))) != cudaSuccess)
                return e;
        }
        if (fuse != nullptr && (e = stage_fuse(*s, stream, *fuse, a - first, p.samples, grid)) != cudaSuccess)
            return e;
#define PTB_LAUNCH_MAT(NSV, GENV, ...)                                                                                                       \
    trace_materials_kernel<NSV, GENV, ##__VA_ARGS__><<<grid, kTraceThreads, smem, stream>>>(pl, spheres, mm, mp.max_depth, mp.rr_start,          \
                                                                             p.sphere_count,                                                  \
                                                                             p.sphere_stride, mp.hit_epsilon, 1.0f, mp.seed, pp, stats,        \
                                                                             s->work_counter)
        if (use_tree) {
            // chunks of consecutive paths per warp claim: 32..64 paths (one or two shading rounds); small, because the last
            // chunk of the unluckiest warp is the kernel's tail and a path here can run to 64 bounces
            int64_t chunk = m / (static_cast<int64_t>(grid) * kWarpsPerBlock * 8);
            chunk = (chunk + 31) / 32 * 32;
            chunk = chunk < 32 ? 32 : (chunk > PTB_BVH_MAX_CHUNK ? PTB_BVH_MAX_CHUNK : chunk);
            if (gen)
                trace_materials_bvh_kernel<true><<<grid, kTraceThreads, smem, stream>>>(pl, mm, mp.max_depth, mp.rr_start, p.sphere_count,
                                                                                      mp.hit_epsilon, 1.0f, mp.seed, pp, stats, bvh, s->work_counter,
                                                                                      static_cast<unsigned int>(chunk));
            else
                trace_materials_bvh_kernel<false><<<grid, kTraceThreads, smem, stream>>>(pl, mm, mp.max_depth, mp.rr_start, p.sphere_count,
                                                                                       mp.hit_epsilon, 1.0f, mp.seed, pp, stats, bvh, s->work_counter,
                                                                                       static_cast<unsigned int>(chunk));
        } else if (ten) {
            if (fuse) PTB_LAUNCH_MAT(10, true, true); else if (gen) PTB_LAUNCH_MAT(10, true); else PTB_LAUNCH_MAT(10, false);
        } else {
            if (fuse) PTB_LAUNCH_MAT(0, true, true); else if (gen) PTB_LAUNCH_MAT(0, true); else PTB_LAUNCH_MAT(0, false);
        }
#undef PTB_LAUNCH_MAT
        if ((e = cudaGetLastError()) != cudaSuccess)
            return e;
    }
    if ((e = cudaEventRecord(s->scene_free, stream)) != cudaSuccess)
        return e;
    s->last_stream = stream;
    s->have_last = true;
    return cudaSuccess;
}

}  // namespace ptb200
