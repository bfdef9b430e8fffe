// mrs_kernels.cu -- sm_100a kernels + C ABI of the B200-native mrs-gym step path.
//
// Two execution shapes, chosen per call from N (agents per env):
//
//  * N <= 32  "group" path: one lane per agent, an env is a group of G = pow2ceil(N) lanes of one
//    warp (N=8: 4 envs per warp).  Everything of MRS.step that touches the simulator is ONE kernel:
//    action -> cascaded PID / mixer -> rotor wrench + ground effect + drag + downwash ->
//    Bullet velocity update -> sphere / ground contact -> position + quaternion integration ->
//    newest X slice + newest A slice.  The three intra-env pair passes (downwash on pre-step
//    positions, contact on unconstrained velocities, adjacency on post-step positions) go through a
//    per-warp shared-memory tile read with 128-bit LDS.  With T > 1 (mrs_step_many) the state stays
//    in registers across steps, only actions stream in and X/A stream out.
//
//  * N > 32   "tiled" path: an env spans many CTAs, so the step is three kernels with the pair
//    passes tiled n-body style through shared memory:  pre (forces -> v*, w*; copies pre-step
//    positions to scratch), post (contact, integration, X), adjacency (row tile x column tile,
//    coalesced 128-bit stores; this is a pure streaming store at N=4096).
//
// Reference behaviour: /root/reference/mrsgym/MRS.py:240-277 and callees (see mrs_device.cuh);
// Bullet step restated in oracle/bullet_model.py.  No CPU fallback exists in this library.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mrs_b200.h"
#include "mrs_device.cuh"
#include "mrs_baked.cuh"

namespace mrs {

constexpr int kBlock = 128;
constexpr unsigned kFull32 = 0xffffffffu;

struct StepArgs {
    const float* actions;  // [T][E][N][A]
    int T;
    int slot_x, slot_a;    // step t writes X tape slot slot_x - t and A tape slot slot_a - t
    // group path: the same as pointers, resolved once on the host (NULL = that tape is not written):
    // step t writes X to X0 - t * xstride and A to A0 - t * astride (strides in floats)
    float* X0;
    float* A0;
    long long xstride, astride;
    int G;                 // group width (power of two >= N), group path only
    int chunk_lo, nchunks; // this launch walks the warp-sized work items [chunk_lo, nchunks), group path only
};

// ------------------------------------------------------------------------------ state planes
__device__ __forceinline__ void load_agent(const float* __restrict__ st, unsigned S, unsigned s, Agent& a) {
    a.px = st[0 * S + s]; a.py = st[1 * S + s]; a.pz = st[2 * S + s];
    a.qx = st[3 * S + s]; a.qy = st[4 * S + s]; a.qz = st[5 * S + s]; a.qw = st[6 * S + s];
    a.vx = st[7 * S + s]; a.vy = st[8 * S + s]; a.vz = st[9 * S + s];
    a.wx = st[10 * S + s]; a.wy = st[11 * S + s]; a.wz = st[12 * S + s];
}

__device__ __forceinline__ void store_agent(float* __restrict__ st, unsigned S, unsigned s, const Agent& a) {
    st[0 * S + s] = a.px; st[1 * S + s] = a.py; st[2 * S + s] = a.pz;
    st[3 * S + s] = a.qx; st[4 * S + s] = a.qy; st[5 * S + s] = a.qz; st[6 * S + s] = a.qw;
    st[7 * S + s] = a.vx; st[8 * S + s] = a.vy; st[9 * S + s] = a.vz;
    st[10 * S + s] = a.wx; st[11 * S + s] = a.wy; st[12 * S + s] = a.wz;
}

__device__ __forceinline__ void dummy_agent(Agent& a) {
    a.px = a.py = 0.f; a.pz = 1.0e3f;
    a.qx = a.qy = a.qz = 0.f; a.qw = 1.f;
    a.vx = a.vy = a.vz = 0.f;
    a.wx = a.wy = a.wz = 0.f;
}

template <int MODE>
__device__ __forceinline__ void load_ctrl(const float* __restrict__ ct, unsigned S, unsigned s, Ctrl& k) {
    using MT = ModeTraits<MODE>;
    if constexpr (MT::io) {
#pragma unroll
        for (int i = 0; i < 3; ++i) k.io[i] = ct[(0 + i) * S + s];
    }
    if constexpr (MT::ip) {
#pragma unroll
        for (int i = 0; i < 3; ++i) k.ip[i] = ct[(3 + i) * S + s];
    }
    if constexpr (MT::vel) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            k.iv[i] = ct[(6 + i) * S + s];
            k.lve[i] = ct[(9 + i) * S + s];
            k.dve[i] = ct[(12 + i) * S + s];
            k.ltv[i] = ct[(15 + i) * S + s];
        }
    }
}

template <int MODE>
__device__ __forceinline__ void store_ctrl(float* __restrict__ ct, unsigned S, unsigned s, const Ctrl& k) {
    using MT = ModeTraits<MODE>;
    if constexpr (MT::io) {
#pragma unroll
        for (int i = 0; i < 3; ++i) ct[(0 + i) * S + s] = k.io[i];
    }
    if constexpr (MT::ip) {
#pragma unroll
        for (int i = 0; i < 3; ++i) ct[(3 + i) * S + s] = k.ip[i];
    }
    if constexpr (MT::vel) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            ct[(6 + i) * S + s] = k.iv[i];
            ct[(9 + i) * S + s] = k.lve[i];
            ct[(12 + i) * S + s] = k.dve[i];
            ct[(15 + i) * S + s] = k.ltv[i];
        }
    }
}

template <int MODE>
__device__ __forceinline__ bool load_action(const float* __restrict__ actions, size_t idx, float* act) {
    constexpr int A = ModeTraits<MODE>::A;
    if constexpr (A == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(actions) + idx);
        act[0] = v.x; act[1] = v.y; act[2] = v.z; act[3] = v.w;
        return isnan(v.x) || isnan(v.y) || isnan(v.z) || isnan(v.w);
    } else if constexpr (A == 3) {
        const float* p = actions + idx * 3;
        act[0] = __ldg(p); act[1] = __ldg(p + 1); act[2] = __ldg(p + 2); act[3] = 0.f;
        return isnan(act[0]) || isnan(act[1]) || isnan(act[2]);
    } else {
        act[0] = act[1] = act[2] = act[3] = 0.f;
        return false;
    }
}

// tape stores are write-once streams (nobody on the device re-reads a slot soon): streaming hint
#ifndef MRS_TAPE_STREAMING
#define MRS_TAPE_STREAMING 1
#endif
#if MRS_TAPE_STREAMING
#define MRS_TAPE_ST(ptr, val) __stcs((ptr), (val))
#else
#define MRS_TAPE_ST(ptr, val) (*(ptr) = (val))
#endif

// newest X slice of one agent (Environment.get_X with the built-in state_fn layouts)
__device__ __forceinline__ void write_X(float* __restrict__ Xs, int layout, unsigned s, const Agent& a) {
    if (layout == MRS_X_POS_VEL) {
        float2* p = reinterpret_cast<float2*>(Xs + (size_t)s * 6);
        MRS_TAPE_ST(p + 0, make_float2(a.px, a.py));
        MRS_TAPE_ST(p + 1, make_float2(a.pz, a.vx));
        MRS_TAPE_ST(p + 2, make_float2(a.vy, a.vz));
    } else if (layout == MRS_X_FULL) {
        float* p = Xs + (size_t)s * 13;
        p[0] = a.px; p[1] = a.py; p[2] = a.pz;
        p[3] = a.qx; p[4] = a.qy; p[5] = a.qz; p[6] = a.qw;
        p[7] = a.vx; p[8] = a.vy; p[9] = a.vz;
        p[10] = a.wx; p[11] = a.wy; p[12] = a.wz;
    }
}

// ------------------------------------------------------------------------------ async staging
// The group kernel walks several warp-chunks per warp.  While chunk c is being computed, the 13
// state planes (+ the first action) of chunk c+1 are already in flight to a per-warp shared-memory
// stage through cp.async (LDGSTS): the load latency of a chunk is hidden behind the arithmetic of
// the previous one without holding a second register copy of the state.  Stage layout per warp
// (floats): 13 state planes x 32 | the mode's PID planes x 32 | 32 actions x ACTION_DIM; everything is
// moved as 16-byte pieces (a plane's share of a chunk is 128 contiguous bytes = 8 pieces).
#ifndef MRS_PREFETCH
#define MRS_PREFETCH 1
#endif

__device__ __forceinline__ void cp_async4(float* smem, const float* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem, const float* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------ group path
// GT > 0: compile-time group width with N == GT (8, 16, 32: pair loops unrolled) and FULL chunks only:
//         every chunk is 32 consecutive valid agent slots, so the kernel carries no bounds logic at
//         all; a ragged last chunk (E not a multiple of 32/N) is a second, one-chunk launch of the
//         GT == 0 kernel (dispatch_step).
// GT == 0: run-time width a.G >= N (any N <= 32), lanes >= N of a group idle.
// WPB = warps per CTA.  WPB = 4: many small CTAs, chunks handed out grid-stride (small jobs).
// WPB = 4 * minb (one CTA owns a whole SM): the CTA takes a contiguous share of the chunks and its
// warps pull them from a shared-memory counter.  Per-warp timestamps at C5 (tools/trace_c5.py)
// showed that with a static split the warps of one launch finish between 12 and 21 us after the
// start -- the hardware warp scheduler is not fair -- and the kernel lasts as long as its slowest
// warp; the SM-local dynamic hand-out keeps all warps of an SM busy until its share is done
// (first / last warp end 14.7 / 20.2 us).  A device-wide atomic counter was tried first: 12k
// same-address L2 atomics per launch serialise and cost +40 %.
//
// BAKED: the model constants (cf2x.urdf, QuadControl gains, Bullet defaults, DT, GRAVITY and what the
// host derives from them) are compile-time values taken from the generated mrs_baked.cuh instead of
// kernel parameters.  sm_100a has no constant-bank operands on its FP instructions: every parameter a
// chunk uses costs an LDC/LDCU issue slot (~10 % of the generic kernel's instructions), immediates cost
// nothing and fold.  The host picks the baked kernel only when the caller's MrsConfig carries exactly
// those values (config_is_baked: bitwise compare); anything else runs the generic kernel.

// plane `pl` of an SoA buffer given the pointer to the agent's slot in plane 0: one IMAD.WIDE.U32
__device__ __forceinline__ float* plane_ptr(float* p0, unsigned S, int pl) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(p0) + (unsigned long long)S * (unsigned)(pl * 4));
}
__device__ __forceinline__ const float* plane_ptr(const float* p0, unsigned S, int pl) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(p0) + (unsigned long long)S * (unsigned)(pl * 4));
}

// dynamic shared memory of one warp of the group kernel: pair tile (+ prefetch stage for full chunks)
template <int MODE, int GT> __host__ __device__ constexpr int group_warp_smem_bytes() {
    return 1024 + ((MRS_PREFETCH && GT != 0) ? mode_stage_floats<MODE>() * 4 : 0);
}

template <int MODE, int GT, int WPB, bool BAKED>
__global__ void __launch_bounds__(WPB * 32, WPB == 4 ? ModeTraits<MODE>::minb : 1)
step_group_kernel(const __grid_constant__ MrsConfig c_in, const __grid_constant__ Derived d_in, const MrsBuffers b,
                  const StepArgs a) {
    MrsConfig c_bk;
    Derived d_bk;
    if constexpr (BAKED) baked_fill(c_bk, d_bk, c_in, d_in);
    const MrsConfig& c = BAKED ? c_bk : c_in;
    const Derived& d = BAKED ? d_bk : d_in;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool kFull = GT != 0;
    constexpr bool kStage = MRS_PREFETCH && kFull;
    // shared memory, one contiguous region per warp so that every address is one per-warp base plus an
    // immediate: [32 positions | 32 velocities] (the pair tile, 1 KB) [prefetch stage]
    constexpr int kWarpBytes = group_warp_smem_bytes<MODE, GT>();
    __shared__ int sh_counter, sh_hi;
    __shared__ unsigned sh_events[5];       // CTA-level status word + the four statistics counters
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + wib * kWarpBytes;
    float4* wpos = reinterpret_cast<float4*>(wbase);
    float4* wvel = wpos + 32;
    float* stage = reinterpret_cast<float*>(wbase + 1024);
    const int G = GT ? GT : a.G;
    const int N = GT ? GT : c.N;
    const int E = c.E;
    const int gpw = 32 / G;
    const int ai = lane & (G - 1);
    const int gb = lane - ai;
    const unsigned S = (unsigned)E * (unsigned)N;
    const int wtotal = gridDim.x * WPB;
    const MrsPhysicsParams& ph = c.phys;
    const bool pair_contact = ph.agent_contact && N > 1;

    // Work distribution (see the comment above the kernel) over the chunks [a.chunk_lo, a.nchunks).
    // kLocal: the CTA owns the contiguous share [cta_lo, cta_hi) and its warps draw chunk indices from
    // a shared-memory counter; a warp always knows its next chunk (the stage prefetch needs it) and
    // draws the one after next at the top of an iteration, so the atomic's latency is never waited for.
    constexpr bool kLocal = WPB > 4;
    const int gw = blockIdx.x * WPB + wib;
    if (threadIdx.x < 5) sh_events[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        const int nwork = a.nchunks - a.chunk_lo;
        sh_counter = a.chunk_lo + (int)(((long long)blockIdx.x * nwork) / gridDim.x);
        sh_hi = a.chunk_lo + (int)(((long long)(blockIdx.x + 1) * nwork) / gridDim.x);
    }
    __syncthreads();
    // lane 0 draws a chunk index (-1 when the share is used up); the others get it by shuffle later
    auto draw = [&]() -> int {
        int v = -1;
        if (lane == 0) {
            // .inc with the maximal bound == add 1; ptxas wraps an .add of a constant in its warp-aggregation
            // sequence (vote, popc, ltmask, shuffle: 14 instructions) although only one lane is active here
            // (ptxas wraps this in its warp-aggregation sequence -- vote, popc, ltmask, shuffle -- although one
            // lane is active; .inc, a run-time addend and an addend read from shared memory all end up the same)
            asm volatile("atom.shared.add.u32 %0, [%1], 1;"
                         : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sh_counter)) : "memory");
            v = (v < sh_hi) ? v : -1;
        }
        return v;
    };
    // chunk = 32 consecutive agent slots (kFull).  The stage is filled in 16-byte pieces, piece q at float
    // offset 4 q, lane l moves pieces l, l + 32, ...:  [0, 96) state planes 0-11 (plane q >> 3, sub-piece
    // q & 7: a lane's plane advances by 4 per round, so its element index is lane part + round part +
    // chunk part), then the chunk's actions (32 * ACTION_DIM contiguous floats: one whole round for
    // ACTION_DIM 4), then plane 12, then 8 pieces per PID plane.
    constexpr int kNC = mode_nctrl<MODE>();
    constexpr int kA = ModeTraits<MODE>::A;
    constexpr int kActBeg = 96, kActEnd = kActBeg + 8 * kA;       // piece ranges
    constexpr int kP12Beg = kActEnd, kP12End = kP12Beg + 8;
    constexpr int kCtlBeg = kP12End, kCtlEnd = kCtlBeg + 8 * kNC;
    constexpr int kPieces = kCtlEnd;
    static_assert(kPieces * 4 == mode_stage_floats<MODE>(), "stage layout");
    const unsigned lane_el = (unsigned)(lane >> 3) * S + (unsigned)(lane & 7) * 4u;   // plane (l>>3), sub-piece (l&7)
    auto prefetch = [&](int chunk) {
        const unsigned s0 = (unsigned)chunk * 32u;
#pragma unroll
        for (int i = 0; i < (kPieces + 31) / 32; ++i) {
            const int q = lane + 32 * i;
            const int lo = 32 * i, hi = 32 * i + 31;       // compile-time after unrolling: the tests below fold
            if (hi < kActBeg) {
                cp_async16(stage + 4 * q, b.state + (lane_el + s0 + (unsigned)(4 * i) * S));
            } else {
                if (lo < kActEnd && hi >= kActBeg && (lo >= kActBeg || q >= kActBeg) && (hi < kActEnd || q < kActEnd))
                    cp_async16(stage + 4 * q, a.actions + (s0 * (unsigned)kA + (unsigned)(q - kActBeg) * 4u));
                if (lo < kP12End && hi >= kP12Beg && (lo >= kP12Beg || q >= kP12Beg) && (hi < kP12End || q < kP12End))
                    cp_async16(stage + 4 * q, b.state + (12u * S + s0 + (unsigned)(q - kP12Beg) * 4u));
                if (kNC > 0 && lo < kCtlEnd && hi >= kCtlBeg && (lo >= kCtlBeg || q >= kCtlBeg) && (hi < kCtlEnd || q < kCtlEnd)) {
                    const int cq = q - kCtlBeg;
                    cp_async16(stage + 4 * q, b.ctrl + ((unsigned)mode_ctrl_plane<MODE>(cq >> 3) * S + s0 + (unsigned)(cq & 7) * 4u));
                }
            }
        }
        cp_async_commit();
    };
#ifdef MRS_TRACE
    // debug build (tools/trace_c5.py): per-warp timeline (globaltimer ns) into bufs.scratch seen as
    // u64[warps][8], plus per-step aggregates [min start, max wait release, max end, min end]
    unsigned long long* trace = reinterpret_cast<unsigned long long*>(b.scratch) + (size_t)gw * 8;
    unsigned long long* agg = reinterpret_cast<unsigned long long*>(b.scratch) + (size_t)8192 * 8 + (size_t)a.slot_x * 4;
    int trace_i = 0;
    auto stamp = [&]() {
        unsigned long long tns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
        if (lane == 0 && trace_i < 8) trace[trace_i] = tns;
        if (lane == 0 && trace_i == 0) atomicMin(agg + 0, tns);
        if (lane == 0 && trace_i == 1) atomicMax(agg + 1, tns);
        ++trace_i;
    };
    auto stamp_end = [&]() {
        unsigned long long tns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
        if (lane == 0) { atomicMax(agg + 2, tns); atomicMin(agg + 3, tns); }
    };
    stamp();
#endif
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef MRS_TRACE
    stamp();
#endif
    int chunk, chunk_next;
    if (kLocal) {
        chunk = __shfl_sync(kFull32, draw(), 0);
        chunk_next = __shfl_sync(kFull32, chunk >= 0 ? draw() : -1, 0);
    } else {
        chunk = a.chunk_lo + gw;
        chunk_next = chunk + wtotal;
        if (chunk >= a.nchunks) chunk = -1;
        if (chunk_next >= a.nchunks) chunk_next = -1;
    }
    if (kStage && chunk >= 0) prefetch(chunk);
    while (chunk >= 0) {
        const int ticket = (kLocal && chunk_next >= 0) ? draw() : -1;   // the chunk after next; warp-uniform condition
        // kFull: the chunk is 32 consecutive valid slots; else lanes >= N of a group (and envs >= E) idle
        const int e = chunk * gpw + (lane / G);
        const bool valid = kFull ? true : ((e < E) && (ai < N));
        const unsigned s = kFull ? (unsigned)chunk * 32u + (unsigned)lane
                                 : (valid ? (unsigned)e * (unsigned)N + (unsigned)ai : 0u);
        Agent st;
        Ctrl k;
        float4 act0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (kStage) {
            cp_async_wait_all();
            __syncwarp();           // pieces were fetched by other lanes
            st.px = stage[0 * 32 + lane]; st.py = stage[1 * 32 + lane]; st.pz = stage[2 * 32 + lane];
            st.qx = stage[3 * 32 + lane]; st.qy = stage[4 * 32 + lane]; st.qz = stage[5 * 32 + lane];
            st.qw = stage[6 * 32 + lane];
            st.vx = stage[7 * 32 + lane]; st.vy = stage[8 * 32 + lane]; st.vz = stage[9 * 32 + lane];
            st.wx = stage[10 * 32 + lane]; st.wy = stage[11 * 32 + lane]; st.wz = stage[4 * kP12Beg + lane];
            const float* cs = stage + 4 * kCtlBeg + lane;
            if constexpr (ModeTraits<MODE>::io) {
#pragma unroll
                for (int i = 0; i < 3; ++i) k.io[i] = cs[i * 32];
            }
            if constexpr (ModeTraits<MODE>::ip) {
#pragma unroll
                for (int i = 0; i < 3; ++i) k.ip[i] = cs[(3 + i) * 32];
            }
            if constexpr (ModeTraits<MODE>::vel) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    k.iv[i] = cs[(3 + i) * 32]; k.lve[i] = cs[(6 + i) * 32];
                    k.dve[i] = cs[(9 + i) * 32]; k.ltv[i] = cs[(12 + i) * 32];
                }
            }
            const float* as = stage + 4 * kActBeg + kA * lane;
            if constexpr (kA == 4) act0 = *reinterpret_cast<const float4*>(as);
            if constexpr (kA == 3) act0 = make_float4(as[0], as[1], as[2], 0.f);
            __syncwarp();           // everyone has read its column before the stage is refilled
            if (chunk_next >= 0) prefetch(chunk_next);
        } else if (valid) {
            load_agent(b.state, S, s, st);
            load_ctrl<MODE>(b.ctrl, S, s, k);
        } else {
            dummy_agent(st);
#pragma unroll
            for (int i = 0; i < 3; ++i) k.io[i] = k.ip[i] = k.iv[i] = k.lve[i] = k.dve[i] = k.ltv[i] = 0.f;
        }
        float* Xs = a.X0;       // tape slots of step t (they move down one slot per step)
        float* As = a.A0;
        for (int t = 0; t < a.T; ++t, Xs -= a.xstride, As -= a.astride) {
            // per-step event word: the registers behind it live only as long as the step needs them
            unsigned status = 0;
            unsigned n_agent_rows = 0, n_ground = 0;
            float rpm[4];
            float act[4];
            bool nan_act = false;
            if (kStage && t == 0) {
                act[0] = act0.x; act[1] = act0.y; act[2] = act0.z; act[3] = act0.w;
                nan_act = kA > 0 && (isnan(act0.x) || isnan(act0.y) || isnan(act0.z) || isnan(act0.w));
            } else if (valid) {
                nan_act = load_action<MODE>(a.actions, (size_t)t * S + s, act);
            } else {
                act[0] = act[1] = act[2] = act[3] = 0.f;
            }
            if (nan_act) status |= MRS_STATUS_NAN_ACTION;

#ifdef MRS_EXP_COPYONLY      // experiment (profiles/README.md): memory movement of a step only, no physics
            st.px += act[0] * 1e-12f;
            if (false) {
#else
            {
#endif
            float R[9];
            quat_to_mat(st, R);
            action_to_rpm<MODE>(c, c_in.quad, d, st, R, act, k, rpm);
            if (b.rpm && MODE != MRS_NO_ACTION && valid) {       // optional Quadcopter.speeds mirror
#pragma unroll
                for (int i = 0; i < 4; ++i) *plane_ptr(b.rpm + s, S, i) = rpm[i];
            }

            // ---- pair pass 1: downwash + contact proximity on the pre-step positions
            __syncwarp();
            wpos[lane] = make_float4(st.px, st.py, st.pz, 0.f);
            __syncwarp();
            float dw = 0.f;
            bool near = false;
            if (MODE != MRS_NO_ACTION || pair_contact) {
#pragma unroll 8
                for (int r = 1; r < G; ++r) {
                    const int j = ai ^ r;
                    if (GT || j < N) {
                        const float4 pj = wpos[gb + j];
                        const float rx = pj.x - st.px, ry = pj.y - st.py, rz = pj.z - st.pz;
                        const float dxy2 = rx * rx + ry * ry;
                        if (MODE != MRS_NO_ACTION) dw += downwash_pair(c.quad, d, dxy2, rz);
                        near = near || (dxy2 + rz * rz < d.lim2);
                    }
                }
            }
            apply_wrench<MODE != MRS_NO_ACTION>(c, d, st, R, rpm, dw);

            // ---- pair pass 2 (rare): sphere-sphere contact on the unconstrained velocities
            if (pair_contact && __any_sync(kFull32, near && valid)) {
                wvel[lane] = make_float4(st.vx, st.vy, st.vz, 0.f);
                __syncwarp();
                if (near) {
                    // st.p still is the pre-step position (integrate comes later)
                    const float p0x = st.px, p0y = st.py, p0z = st.pz;
                    float acc[3] = {0.f, 0.f, 0.f};
                    for (int r = 1; r < G; ++r) {
                        const int j = ai ^ r;
                        if (GT || j < N) {
                            const float4 pj = wpos[gb + j];
                            const float4 vj = wvel[gb + j];
                            if (agent_contact_pair(ph, d, p0x - pj.x, p0y - pj.y, p0z - pj.z, st.vx - vj.x,
                                                   st.vy - vj.y, st.vz - vj.z, acc))
                                ++n_agent_rows;
                        }
                    }
                    st.vx += acc[0]; st.vy += acc[1]; st.vz += acc[2];
                }
            }
            if (ph.ground_contact && ground_contact(ph, d, st)) ++n_ground;
            integrate(c, d, st);
            if (!agent_finite(st)) status |= MRS_STATUS_NONFINITE;

            }
            // ---- observation: newest X slice and newest A slice into their tape slots
            // (X rows staged through shared memory and written as whole 16-byte pieces were measured: neutral
            // at C5, 18.4 vs 18.1 us -- unlike the A rows below, the float2 stores are not the limiter)
            if (a.X0 && valid) write_X(Xs, c.state_layout, s, st);
            if (a.A0) {
                float* Arow = As + (size_t)s * N;
                if (d.comm_inf) {
                    if (valid)
                        for (int j = 0; j < N; ++j) Arow[j] = (j == ai) ? 0.f : 1.f;
                } else {
                    __syncwarp();
                    wpos[lane] = make_float4(st.px, st.py, st.pz, 0.f);
                    __syncwarp();
                    if constexpr (GT != 0) {
                        // Lane pair (2k, 2k+1) shares the two rows 2k, 2k+1 of A: the even lane computes the
                        // even column quads of BOTH rows, the odd lane the odd quads.  One STG.128 of the pair
                        // then covers 32 contiguous bytes, i.e. whole 32-byte sectors (a lane writing its own
                        // row alone sends every sector twice, half filled), and every column position read
                        // from shared memory serves two rows.
                        const int h = lane & 1;
                        const float4 r0 = wpos[lane & ~1], r1 = wpos[lane | 1];
                        float* Arow0 = As + (size_t)(s & ~1u) * GT;
                        const int d0 = (ai & ~1) - 4 * h;          // column of row 0's diagonal relative to quad h
#pragma unroll
                        for (int qq = 0; qq < GT / 8; ++qq) {       // this lane's quads: h, h + 2, ...
                            const int col = 8 * qq + 4 * h;
                            float4 pj[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) pj[u] = wpos[gb + col + u];
                            float v0[4], v1[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float h0 = adjacency_pair(r0.x, r0.y, r0.z, pj[u].x, pj[u].y, pj[u].z, d.s_max);
                                const float h1 = adjacency_pair(r1.x, r1.y, r1.z, pj[u].x, pj[u].y, pj[u].z, d.s_max);
                                v0[u] = (8 * qq + u == d0) ? 0.f : h0;
                                v1[u] = (8 * qq + u == d0 + 1) ? 0.f : h1;
                            }
                            MRS_TAPE_ST(reinterpret_cast<float4*>(Arow0 + col), make_float4(v0[0], v0[1], v0[2], v0[3]));
                            MRS_TAPE_ST(reinterpret_cast<float4*>(Arow0 + GT + col), make_float4(v1[0], v1[1], v1[2], v1[3]));
                        }
                    } else if (valid) {
                        if ((N & 3) == 0) {
#pragma unroll 2
                            for (int j = 0; j < N; j += 4) {
                                // loads and arithmetic unconditional, the diagonal is a select afterwards: a
                                // conditional around the shared-memory load compiles to one divergent
                                // BSSY/BRA/BSYNC block per element (no overlap between the pairs)
                                float4 pj[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) pj[u] = wpos[gb + j + u];
                                float v[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float hit = adjacency_pair(st.px, st.py, st.pz, pj[u].x, pj[u].y, pj[u].z, d.s_max);
                                    v[u] = (j + u == ai) ? 0.f : hit;
                                }
                                MRS_TAPE_ST(reinterpret_cast<float4*>(Arow + j), make_float4(v[0], v[1], v[2], v[3]));
                            }
                        } else {
                            for (int j = 0; j < N; ++j) {
                                const float4 pj = wpos[gb + j];
                                const float hit = adjacency_pair(st.px, st.py, st.pz, pj.x, pj.y, pj.z, d.s_max);
                                Arow[j] = (j == ai) ? 0.f : hit;
                            }
                        }
                    }
                }
            }
            // status / statistics: warp-reduce, then CTA-level shared-memory counters; the global atomics
            // happen once per CTA at the end.  (A swarm resting on the ground reports a ground contact per
            // agent per step: with one global atomic per warp-chunk that was 10 k same-address L2 atomics
            // per launch at C5 and cost ~15 % of the step.)
            if (!valid) { status = 0; n_agent_rows = 0; n_ground = 0; }
            if (__reduce_or_sync(kFull32, status | n_agent_rows | n_ground)) {      // rare in free flight
                const unsigned any_status = __reduce_or_sync(kFull32, status);
                const unsigned sum_rows = __reduce_add_sync(kFull32, n_agent_rows);
                const unsigned sum_gnd = __reduce_add_sync(kFull32, n_ground);
                if (lane == 0) {
                    if (any_status) atomicOr(&sh_events[0], any_status);
                    if (sum_rows) atomicAdd(&sh_events[1], sum_rows);
                    if (sum_gnd) atomicAdd(&sh_events[2], sum_gnd);
                    if (any_status & MRS_STATUS_NONFINITE) atomicAdd(&sh_events[3], 1u);
                    if (any_status & MRS_STATUS_NAN_ACTION) atomicAdd(&sh_events[4], 1u);
                }
            }
        }

        if (valid) {
            float* p0 = b.state + s;
            *plane_ptr(p0, S, 0) = st.px; *plane_ptr(p0, S, 1) = st.py; *plane_ptr(p0, S, 2) = st.pz;
            *plane_ptr(p0, S, 3) = st.qx; *plane_ptr(p0, S, 4) = st.qy; *plane_ptr(p0, S, 5) = st.qz;
            *plane_ptr(p0, S, 6) = st.qw;
            *plane_ptr(p0, S, 7) = st.vx; *plane_ptr(p0, S, 8) = st.vy; *plane_ptr(p0, S, 9) = st.vz;
            *plane_ptr(p0, S, 10) = st.wx; *plane_ptr(p0, S, 11) = st.wy; *plane_ptr(p0, S, 12) = st.wz;
            store_ctrl<MODE>(b.ctrl, S, s, k);
        }
#ifdef MRS_TRACE
        stamp();
#endif
        chunk = chunk_next;
        if (kLocal) {
            chunk_next = __shfl_sync(kFull32, ticket, 0);
        } else {
            chunk_next = (chunk >= 0 && chunk + wtotal < a.nchunks) ? chunk + wtotal : -1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sh_events[0] && b.status) atomicOr(b.status, sh_events[0]);
        if (b.stats) {
            if (sh_events[1]) atomicAdd(b.stats + MRS_STAT_AGENT_CONTACTS, (unsigned long long)sh_events[1]);
            if (sh_events[2]) atomicAdd(b.stats + MRS_STAT_GROUND_CONTACTS, (unsigned long long)sh_events[2]);
            if (sh_events[3]) atomicAdd(b.stats + MRS_STAT_NONFINITE, (unsigned long long)sh_events[3]);
            if (sh_events[4]) atomicAdd(b.stats + MRS_STAT_NAN_ACTIONS, (unsigned long long)sh_events[4]);
        }
    }
#ifdef MRS_TRACE
    stamp_end();
#endif
}

// ------------------------------------------------------------------------------ wide path (N > 32)
// An env no longer fits a warp, so the pair passes are spread over LPA lanes PER AGENT (LPA = 8 for
// N <= 128, else 32): every lane walks the partners j = l, l + LPA, ... of its agent straight from
// the L1/L2-resident position planes, the partial sums are combined with a fixed-order xor-shuffle
// tree (deterministic), and the group's lane 0 runs the per-agent part.  One env of 4096 agents
// therefore fills the GPU with 4096 warps instead of 32 CTAs.
// scratch planes: 0-2 unconstrained velocity, 3-5 pre-step position, 6 contact-proximity flag.
template <int LPA>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPA / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull32, v, o);
    return v;
}

// per-agent part of the wide pre pass: controller -> rotor wrench + aero (dw = downwash sum) ->
// unconstrained velocities; stashes v*, the pre-step position and the proximity flag in scratch
template <int MODE>
__device__ __forceinline__ void agent_pre(const MrsConfig& c, const Derived& d, const MrsBuffers& b,
                                          const float* __restrict__ actions, unsigned S, unsigned s, float dw, bool near) {
    Agent st;
    Ctrl k;
    load_agent(b.state, S, s, st);
    load_ctrl<MODE>(b.ctrl, S, s, k);
    const float pix = st.px, piy = st.py, piz = st.pz;
    float act[4];
    unsigned status = 0;
    if (load_action<MODE>(actions, s, act)) status |= MRS_STATUS_NAN_ACTION;
    float R[9], rpm[4];
    quat_to_mat(st, R);
    action_to_rpm<MODE>(c, c.quad, d, st, R, act, k, rpm);
    apply_wrench<MODE != MRS_NO_ACTION>(c, d, st, R, rpm, dw);
    float* sc = b.scratch;
    sc[0 * (size_t)S + s] = st.vx; sc[1 * (size_t)S + s] = st.vy; sc[2 * (size_t)S + s] = st.vz;
    sc[3 * (size_t)S + s] = pix; sc[4 * (size_t)S + s] = piy; sc[5 * (size_t)S + s] = piz;
    sc[6 * (size_t)S + s] = near ? 1.f : 0.f;
    b.state[10 * (size_t)S + s] = st.wx; b.state[11 * (size_t)S + s] = st.wy; b.state[12 * (size_t)S + s] = st.wz;
    store_ctrl<MODE>(b.ctrl, S, s, k);
    if (b.rpm && MODE != MRS_NO_ACTION) {
#pragma unroll
        for (int i = 0; i < 4; ++i) b.rpm[i * (size_t)S + s] = rpm[i];
    }
    if (status && b.status) {
        atomicOr(b.status, status);
        if (b.stats) atomicAdd(b.stats + MRS_STAT_NAN_ACTIONS, 1ull);
    }
}

template <int MODE, int LPA>
__global__ void __launch_bounds__(kBlock)
step_pre_kernel(const __grid_constant__ MrsConfig c, const __grid_constant__ Derived d, const MrsBuffers b,
                const float* __restrict__ actions) {
    const int N = c.N;
    const unsigned S = (unsigned)c.E * (unsigned)N;
    const unsigned gid = (blockIdx.x * kBlock + threadIdx.x) / LPA;      // agent slot of this lane group
    const int l = threadIdx.x & (LPA - 1);
    const bool valid = gid < S;
    const unsigned s = valid ? gid : 0u;
    const unsigned env0 = (s / (unsigned)N) * (unsigned)N;
    const int ai = (int)(s - env0);
    const float* __restrict__ px = b.state + 0 * (size_t)S + env0;
    const float* __restrict__ py = b.state + 1 * (size_t)S + env0;
    const float* __restrict__ pz = b.state + 2 * (size_t)S + env0;
    const float pix = px[ai], piy = py[ai], piz = pz[ai];
    float dw = 0.f;
    bool near = false;
    const bool pair_contact = c.phys.agent_contact && N > 1;
    if (MODE != MRS_NO_ACTION || pair_contact) {
        // Uniform trip count for the whole warp (the vote below needs every lane).  The SFU part of the
        // downwash is skipped for a whole warp when none of its 32 pairs can contribute: partner not
        // above (rz <= 0), dxy >= 10, or exp(-0.5 (dxy/beta)^2) underflowing float32 (0.5 q^2 > 104).
        // Consecutive partners share a height layer in a lattice-like swarm, so the vote is mostly uniform.
        // four partners per lane and iteration: independent loads in flight, one vote per four pairs
        for (int j0 = 0; j0 < N; j0 += 4 * LPA) {
            float dxy2[4], rz[4];
            bool live[4], any_live = false;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * LPA + l;
                const bool in = j < N;
                const int jj = in ? j : ai;
                const float rx = px[jj] - pix, ry = py[jj] - piy;
                rz[u] = pz[jj] - piz;
                dxy2[u] = rx * rx + ry * ry;
                const bool other = in && j != ai;
                const float beta = c.quad.dw2 * rz[u] + c.quad.dw3;
                live[u] = MODE != MRS_NO_ACTION && other && rz[u] > 0.f && dxy2[u] < 100.f &&
                          !(dxy2[u] > 208.f * beta * beta);
                any_live = any_live || live[u];
                near = near || (other && dxy2[u] + rz[u] * rz[u] < d.lim2);
            }
            if (MODE != MRS_NO_ACTION && __any_sync(kFull32, any_live)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float f = downwash_pair(c.quad, d, dxy2[u], rz[u]);
                    dw += live[u] ? f : 0.f;
                }
            }
        }
    }
    static_assert(LPA <= 32, "N > 128 uses pair_tile_kernel");
    dw = group_sum<LPA>(dw);
    const unsigned gmask = (LPA == 32) ? kFull32 : (((1u << (LPA & 31)) - 1u) << ((threadIdx.x & 31) & ~(LPA - 1)));
    near = (__ballot_sync(kFull32, near) & gmask) != 0u;
    if (l != 0 || !valid) return;
    agent_pre<MODE>(c, d, b, actions, S, s, dw, near && pair_contact);
}

// ------------------------------------------------------------------------------ pair pass for N > 128
// n-body tiling.  A CTA owns 128 agents of one env (one per thread, own position in registers) and one
// of `nsplit` slices of the partner range; partner positions go through shared memory in tiles of 128
// and every thread reads the SAME partner (broadcast LDS.128), so a pair costs no global load, no index
// arithmetic and no shuffle: ~16 instructions without the downwash term, which is skipped per warp when
// none of its 32 agents can feel this partner (not above, dxy >= 10 m, or exp(-0.5 (dxy/beta)^2)
// underflowing float32).  Partial sums of slice js go to scratch plane kPairPlane0 + js, the proximity
// flag to plane kPairPlane0 + nsplit + js; agent_pre_kernel adds them in slice order (deterministic).
// (The first version gave every agent a whole CTA that walked the partners from the L1-resident position
// planes: 45 instructions per pair, 29.9 us at N = 4096.)
constexpr int kPairPlane0 = MRS_SCRATCH_PLANES;      // first partial-sum plane
constexpr int kPairMaxSplit = MRS_SCRATCH_PAIR_SPLITS;

template <int MODE>
__global__ void __launch_bounds__(kBlock)
pair_tile_kernel(const __grid_constant__ MrsConfig c, const __grid_constant__ Derived d, const MrsBuffers b, int jw,
                 int nsplit) {
    __shared__ float4 tile[kBlock];
    const int N = c.N;
    const unsigned S = (unsigned)c.E * (unsigned)N;
    const int itiles = (N + kBlock - 1) / kBlock;
    const unsigned per_env = (unsigned)(itiles * nsplit);
    const unsigned env = blockIdx.x / per_env, rem = blockIdx.x - env * per_env;
    const int it = (int)(rem / (unsigned)nsplit), js = (int)(rem - (unsigned)it * (unsigned)nsplit);
    const unsigned env0 = env * (unsigned)N;
    const int i = it * kBlock + threadIdx.x;
    const bool valid = i < N;
    const float* __restrict__ px = b.state + 0 * (size_t)S + env0;
    const float* __restrict__ py = b.state + 1 * (size_t)S + env0;
    const float* __restrict__ pz = b.state + 2 * (size_t)S + env0;
    // an idle lane sits far below everything: no partner is above-and-near, none is close
    const float pix = valid ? px[i] : 0.f, piy = valid ? py[i] : 0.f, piz = valid ? pz[i] : 3.0e18f;
    const bool pair_contact = c.phys.agent_contact && N > 1;
    // 0 < d2 < lim2 as ONE unsigned compare of the float bits (d2 >= 0 or NaN): (bits - 1) < (bits(lim2) - 1);
    // d2 == 0 is the agent itself (or a coincident partner, which the contact row ignores anyway)
    const unsigned lim_m1 = __float_as_uint(d.lim2) - 1u;
    float dw = 0.f;
    bool near = false;
    // Tile culling: a tile whose highest partner lies more than the contact range below the lowest agent of
    // this warp can neither blow on any of them (dz <= 0) nor touch them: skipped as a whole (exact, not a
    // cut-off).  It pays when the index order follows height; the C4 bench lattice has z as its fastest
    // index, so no tile is skipped there and the test costs ~1 %.
    __shared__ float tile_zmax[kBlock / 32];
    float wz_min = valid ? piz : 3.0e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wz_min = fminf(wz_min, __shfl_xor_sync(kFull32, wz_min, o));
    const float z_skip = wz_min - sqrtf(d.lim2);
    const int jbeg = js * jw, jend = min(jbeg + jw, N);
    for (int j0 = jbeg; j0 < jend; j0 += kBlock) {
        const int jj = j0 + threadIdx.x;
        const float4 mine = (jj < jend) ? make_float4(px[jj], py[jj], pz[jj], 0.f) : make_float4(3.0e18f, 0.f, -3.0e18f, 0.f);
        float zm = mine.z;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) zm = fmaxf(zm, __shfl_xor_sync(kFull32, zm, o));
        __syncthreads();
        tile[threadIdx.x] = mine;
        if ((threadIdx.x & 31) == 0) tile_zmax[threadIdx.x >> 5] = zm;
        __syncthreads();
        const float tz = fmaxf(fmaxf(tile_zmax[0], tile_zmax[1]), fmaxf(tile_zmax[2], tile_zmax[3]));
        if (tz < z_skip) continue;          // warp-uniform
#pragma unroll 2
        for (int jl = 0; jl < kBlock; jl += 4) {
            float dxy2[4], rz[4];
            bool live[4], any_live = false;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 pj = tile[jl + u];
                const float rx = pj.x - pix, ry = pj.y - piy;
                rz[u] = pj.z - piz;
                dxy2[u] = rx * rx + ry * ry;
                const float d2 = dxy2[u] + rz[u] * rz[u];
                near = near || (__float_as_uint(d2) - 1u < lim_m1);
                const float beta = c.quad.dw2 * rz[u] + c.quad.dw3;
                live[u] = MODE != MRS_NO_ACTION && rz[u] > 0.f && dxy2[u] < 100.f && !(dxy2[u] > 208.f * beta * beta);
                any_live = any_live || live[u];
            }
            if (MODE != MRS_NO_ACTION && __any_sync(kFull32, any_live)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float f = downwash_pair(c.quad, d, dxy2[u], rz[u]);
                    dw += live[u] ? f : 0.f;
                }
            }
        }
    }
    if (!valid) return;
    const unsigned s = env0 + (unsigned)i;
    b.scratch[(size_t)(kPairPlane0 + js) * S + s] = dw;
    b.scratch[(size_t)(kPairPlane0 + nsplit + js) * S + s] = (near && pair_contact) ? 1.f : 0.f;
}

// thread-per-agent half of the wide pre pass for N >= 1024 (dw and the proximity flag come from scratch)
template <int MODE>
__global__ void __launch_bounds__(kBlock)
agent_pre_kernel(const __grid_constant__ MrsConfig c, const __grid_constant__ Derived d, const MrsBuffers b,
                 const float* __restrict__ actions, int nsplit) {
    const unsigned S = (unsigned)c.E * (unsigned)c.N;
    const unsigned s = blockIdx.x * kBlock + threadIdx.x;
    if (s >= S) return;
    float dw = 0.f, fl = 0.f;
    for (int j0 = 0; j0 < nsplit; j0 += 8) {       // fixed order: the sum does not depend on the launch shape
        float pd[8], pf[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {              // independent loads in flight
            const bool in = j0 + u < nsplit;
            pd[u] = in ? b.scratch[(size_t)(kPairPlane0 + j0 + u) * S + s] : 0.f;
            pf[u] = in ? b.scratch[(size_t)(kPairPlane0 + nsplit + j0 + u) * S + s] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { dw += pd[u]; fl += pf[u]; }
    }
    agent_pre<MODE>(c, d, b, actions, S, s, dw, fl != 0.f);
}

template <int LPA>
__global__ void __launch_bounds__(kBlock)
step_post_kernel(const __grid_constant__ MrsConfig c, const __grid_constant__ Derived d, const MrsBuffers b, int slot) {
    const int N = c.N;
    const unsigned S = (unsigned)c.E * (unsigned)N;
    const unsigned gid = (blockIdx.x * kBlock + threadIdx.x) / LPA;
    const int l = threadIdx.x & (LPA - 1);
    const bool valid = gid < S;
    const unsigned s = valid ? gid : 0u;
    const unsigned env0 = (s / (unsigned)N) * (unsigned)N;
    const int ai = (int)(s - env0);
    const MrsPhysicsParams& ph = c.phys;
    const float* __restrict__ sc = b.scratch;
    Agent st;
    st.px = sc[3 * (size_t)S + s]; st.py = sc[4 * (size_t)S + s]; st.pz = sc[5 * (size_t)S + s];
    st.vx = sc[0 * (size_t)S + s]; st.vy = sc[1 * (size_t)S + s]; st.vz = sc[2 * (size_t)S + s];
    const bool near = valid && sc[6 * (size_t)S + s] != 0.f;       // uniform over the lane group
    float acc[3] = {0.f, 0.f, 0.f};
    unsigned rows = 0;
    if (near) {
        const float* __restrict__ qx = sc + 3 * (size_t)S + env0;
        const float* __restrict__ qy = sc + 4 * (size_t)S + env0;
        const float* __restrict__ qz = sc + 5 * (size_t)S + env0;
        for (int j0 = l; j0 < N; j0 += 4 * LPA) {     // 4 independent partner loads in flight per lane
            float dx[4], dy[4], dz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * LPA;
                const bool ok = j < N && j != ai;
                const int jj = ok ? j : ai;
                dx[u] = st.px - qx[jj]; dy[u] = st.py - qy[jj]; dz[u] = st.pz - qz[jj];
                if (!ok) dx[u] = 1.0e18f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u] < d.lim2) {
                    const int j = j0 + u * LPA;
                    const float vjx = sc[0 * (size_t)S + env0 + j], vjy = sc[1 * (size_t)S + env0 + j],
                                vjz = sc[2 * (size_t)S + env0 + j];
                    if (agent_contact_pair(ph, d, dx[u], dy[u], dz[u], st.vx - vjx, st.vy - vjy, st.vz - vjz, acc)) ++rows;
                }
            }
        }
    }
    // every lane of the warp takes part in the shuffles (groups without contact add zeros)
    acc[0] = group_sum<LPA>(acc[0]); acc[1] = group_sum<LPA>(acc[1]); acc[2] = group_sum<LPA>(acc[2]);
    rows = (unsigned)group_sum<LPA>((float)rows);
    const bool lead = (l == 0) && valid;
    unsigned gnd = 0, bad = 0;
    if (lead) {
        st.vx += acc[0]; st.vy += acc[1]; st.vz += acc[2];
        st.qx = b.state[3 * (size_t)S + s]; st.qy = b.state[4 * (size_t)S + s]; st.qz = b.state[5 * (size_t)S + s];
        st.qw = b.state[6 * (size_t)S + s];
        st.wx = b.state[10 * (size_t)S + s]; st.wy = b.state[11 * (size_t)S + s]; st.wz = b.state[12 * (size_t)S + s];
        if (ph.ground_contact && ground_contact(ph, d, st)) gnd = 1;
        integrate(c, d, st);
        store_agent(b.state, S, s, st);
        if (b.X_tape && c.state_layout != MRS_X_NONE)
            write_X(b.X_tape + (size_t)slot * S * state_dim(c.state_layout), c.state_layout, s, st);
        bad = agent_finite(st) ? 0u : 1u;
    }
    // statistics: one warp reduction, then at most three global atomics per warp (not per agent)
    const unsigned w_rows = __reduce_add_sync(kFull32, lead ? rows : 0u);
    const unsigned w_gnd = __reduce_add_sync(kFull32, gnd);
    const unsigned w_bad = __reduce_add_sync(kFull32, bad);
    if ((threadIdx.x & 31) == 0) {
        if (w_bad && b.status) atomicOr(b.status, MRS_STATUS_NONFINITE);
        if (b.stats) {
            if (w_rows) atomicAdd(b.stats + MRS_STAT_AGENT_CONTACTS, (unsigned long long)w_rows);
            if (w_gnd) atomicAdd(b.stats + MRS_STAT_GROUND_CONTACTS, (unsigned long long)w_gnd);
            if (w_bad) atomicAdd(b.stats + MRS_STAT_NONFINITE, (unsigned long long)w_bad);
        }
    }
}

// ------------------------------------------------------------------------------ adjacency
// Flat version, any N and any position layout: one thread per output element.
// pos component c of agent (e, i) = pos[c * cs + (e * N + i) * as].
__global__ void __launch_bounds__(256)
adjacency_flat_kernel(const float* __restrict__ pos, size_t cs, size_t as, float* __restrict__ A, int E, int N,
                      float s_max, int comm_inf) {
    const size_t total = (size_t)E * N * N;
    for (size_t kx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; kx < total; kx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(kx % N);
        const size_t row = kx / N;
        const int i = (int)(row % N);
        const size_t e = row / N;
        float v;
        if (i == j) v = 0.f;
        else if (comm_inf) v = 1.f;
        else {
            const size_t si = (e * N + i) * as, sj = (e * N + j) * as;
            v = adjacency_pair(pos[si], pos[cs + si], pos[2 * cs + si], pos[sj], pos[cs + sj], pos[2 * cs + sj], s_max);
        }
        A[kx] = v;
    }
}

// Tiled version for N >= 128, N % 4 == 0.  CTA = kRowTile rows x 512 columns of one env; each
// lane keeps its 4 column positions in registers, row positions are broadcast from shared
// memory, each warp store is 512 contiguous bytes of one A row.
constexpr int kRowTile = 32;
__global__ void __launch_bounds__(kBlock)
adjacency_tiled_kernel(const float* __restrict__ pos, size_t cs, size_t as, float* __restrict__ A, int E, int N,
                       float s_max, int comm_inf) {
    __shared__ float4 rows[kRowTile];
    const int col_tiles = (N + 4 * kBlock - 1) / (4 * kBlock);
    const int ct = blockIdx.x % col_tiles;
    const int rt = blockIdx.x / col_tiles;
    const size_t e = blockIdx.y;
    const size_t env0 = e * N;
    const int i0 = rt * kRowTile;
    const int j = ct * 4 * kBlock + threadIdx.x * 4;
    if (threadIdx.x < kRowTile && i0 + threadIdx.x < N) {
        const size_t si = (env0 + i0 + threadIdx.x) * as;
        rows[threadIdx.x] = make_float4(pos[si], pos[cs + si], pos[2 * cs + si], 0.f);
    }
    float xj[4], yj[4], zj[4];
    const bool col_ok = j < N;
    if (col_ok) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t sj = (env0 + j + u) * as;
            xj[u] = pos[sj]; yj[u] = pos[cs + sj]; zj[u] = pos[2 * cs + sj];
        }
    }
    __syncthreads();
    if (!col_ok) return;
    const int nrows = min(kRowTile, N - i0);
    float* out = A + (env0 + i0) * (size_t)N + j;
    for (int r = 0; r < nrows; ++r) {
        const float4 pi = rows[r];
        const int i = i0 + r;
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j + u == i) v[u] = 0.f;
            else if (comm_inf) v[u] = 1.f;
            else v[u] = adjacency_pair(pi.x, pi.y, pi.z, xj[u], yj[u], zj[u], s_max);
        }
        __stcs(reinterpret_cast<float4*>(out + (size_t)r * N), make_float4(v[0], v[1], v[2], v[3]));
    }
}

// X of the current state (MRS.calc_Xk outside step)
__global__ void __launch_bounds__(256)
observe_x_kernel(const MrsBuffers b, unsigned S, int layout, int slot) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    Agent st;
    load_agent(b.state, S, s, st);
    write_X(b.X_tape + (size_t)slot * S * (size_t)state_dim(layout), layout, s, st);
}

// ------------------------------------------------------------------------------ set_state
// Object.set_state (Object.py:42-65): euler 'xyz' (extrinsic) -> quaternion xyzw =
// qz(yaw) * qy(pitch) * qx(roll); NULL component = keep; masked per env.
__global__ void __launch_bounds__(256)
set_state_kernel(float* __restrict__ st, size_t S, int N, const float* __restrict__ pos, const float* __restrict__ ori,
                 const float* __restrict__ vel, const float* __restrict__ angvel, const unsigned char* __restrict__ mask) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (mask && !mask[s / N]) return;
    if (pos) {
        st[0 * S + s] = pos[3 * s]; st[1 * S + s] = pos[3 * s + 1]; st[2 * S + s] = pos[3 * s + 2];
    }
    if (ori) {
        float sr, cr, sp, cp, sy, cy;
        sincosf(0.5f * ori[3 * s], &sr, &cr);
        sincosf(0.5f * ori[3 * s + 1], &sp, &cp);
        sincosf(0.5f * ori[3 * s + 2], &sy, &cy);
        st[3 * S + s] = sr * cp * cy - cr * sp * sy;
        st[4 * S + s] = cr * sp * cy + sr * cp * sy;
        st[5 * S + s] = cr * cp * sy - sr * sp * cy;
        st[6 * S + s] = cr * cp * cy + sr * sp * sy;
    }
    if (vel) {
        st[7 * S + s] = vel[3 * s]; st[8 * S + s] = vel[3 * s + 1]; st[9 * S + s] = vel[3 * s + 2];
    }
    if (angvel) {
        st[10 * S + s] = angvel[3 * s]; st[11 * S + s] = angvel[3 * s + 1]; st[12 * S + s] = angvel[3 * s + 2];
    }
}

// ------------------------------------------------------------------------------ spawn
// On-device MRS.generate_start_pos / generate_start_ori + reset (MRS.py:127-161,174-184) for the
// default spawn distribution (MRS.default_spawn_dist, MRS.py:69-78): z ~ U[z_lo, z_hi], xy ~ N(0,
// sigma) pulled onto the disc of radius xy_radius when outside it (Util.SphereTransform within=True),
// agents closer than 2*AGENT_RADIUS to another agent of their env are re-drawn until none collides
// (the higher-indexed agent of a colliding pair is re-drawn, so at least one of them stays).
// One warp per env (N <= 32), counter-based RNG (splitmix64 of seed / env / agent / draw), so a reset
// is reproducible from (seed, env) alone and independent of the launch shape.
struct SpawnArgs {
    unsigned long long seed;
    float z_lo, z_hi, xy_radius, xy_sigma;
    float yaw_lo, yaw_hi;
    int max_rounds;
};

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ float u01(unsigned long long h) { return ((h >> 40) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(128)
spawn_kernel(const __grid_constant__ MrsConfig c, const MrsBuffers b, const SpawnArgs sp,
             const unsigned char* __restrict__ mask, unsigned* __restrict__ failed) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= c.E) return;
    if (mask && !mask[e]) return;
    const int N = c.N;
    const unsigned S = (unsigned)c.E * (unsigned)N;
    const bool valid = lane < N;
    const float lim2 = 4.f * c.phys.agent_radius * c.phys.agent_radius;
    float x = 0.f, y = 0.f, z = 0.f;
    bool redraw = true;
    int round = 0;
    for (; round < sp.max_rounds; ++round) {
        if (redraw && valid) {
            const unsigned long long key = splitmix64(sp.seed ^ splitmix64(((unsigned long long)e << 20) ^ ((unsigned long long)lane << 12) ^ (unsigned long long)round));
            const float u1 = u01(key), u2 = u01(splitmix64(key)), u3 = u01(splitmix64(key ^ 0x5851F42D4C957F2Dull));
            const float r = sp.xy_sigma * sqrtf(-2.f * logf(u1));          // Box-Muller
            float sn, cs;
            sincosf(6.28318530717958647692f * u2, &sn, &cs);
            x = r * cs; y = r * sn;
            const float mag = fmaxf(sqrtf(x * x + y * y), sp.xy_radius);     // SphereTransform(within=True)
            x = x / mag * sp.xy_radius; y = y / mag * sp.xy_radius;
            z = sp.z_lo + (sp.z_hi - sp.z_lo) * u3;
        }
        bool hit = false;
        for (int j = 0; j < N; ++j) {
            const float xj = __shfl_sync(kFull32, x, j), yj = __shfl_sync(kFull32, y, j), zj = __shfl_sync(kFull32, z, j);
            const float dx = x - xj, dy = y - yj, dz = z - zj;
            hit = hit || (valid && j < lane && dx * dx + dy * dy + dz * dz < lim2);
        }
        redraw = hit;
        if (!__any_sync(kFull32, hit)) break;
    }
    if (round >= sp.max_rounds && lane == 0 && failed) atomicAdd(failed, 1u);
    if (!valid) return;
    const unsigned s = (unsigned)e * (unsigned)N + (unsigned)lane;
    float* st = b.state;
    st[0 * (size_t)S + s] = x; st[1 * (size_t)S + s] = y; st[2 * (size_t)S + s] = z;
    const unsigned long long ky = splitmix64(sp.seed ^ splitmix64(0xA5A5A5A5ull ^ ((unsigned long long)e << 20) ^ ((unsigned long long)lane << 12)));
    const float yaw = sp.yaw_lo + (sp.yaw_hi - sp.yaw_lo) * u01(ky);
    float sy, cy;
    sincosf(0.5f * yaw, &sy, &cy);
    st[3 * (size_t)S + s] = 0.f; st[4 * (size_t)S + s] = 0.f; st[5 * (size_t)S + s] = sy; st[6 * (size_t)S + s] = cy;
#pragma unroll
    for (int p = 7; p < 13; ++p) st[p * (size_t)S + s] = 0.f;
}

// ------------------------------------------------------------------------------ sensors (row f4)
// Analytic sensors of the 'simple' world (ground box top at ground_z, agents as AGENT_RADIUS
// spheres -- the same contact geometry as the step).  One thread per agent (proximity) or per
// (agent, ray) (raycast); partners are walked from the L1/L2-resident position planes.
// Object.collision / get_dist / get_contact_points / raycast (Object.py:100-174) on these primitives.
__global__ void __launch_bounds__(128)
proximity_kernel(const __grid_constant__ MrsConfig c, const MrsBuffers b, float thresh, float* __restrict__ gap_agent,
                 int* __restrict__ nearest, float* __restrict__ gap_ground, unsigned char* __restrict__ collision) {
    const unsigned S = (unsigned)c.E * (unsigned)c.N;
    const unsigned s = blockIdx.x * 128u + threadIdx.x;
    if (s >= S) return;
    const int N = c.N;
    const unsigned env0 = (s / (unsigned)N) * (unsigned)N;
    const int ai = (int)(s - env0);
    const float* px = b.state + env0;
    const float* py = b.state + (size_t)S + env0;
    const float* pz = b.state + 2 * (size_t)S + env0;
    const float x = px[ai], y = py[ai], z = pz[ai];
    float best = INFINITY;
    int arg = -1;
    for (int j = 0; j < N; ++j) {
        if (j == ai) continue;
        const float dx = x - px[j], dy = y - py[j], dz = z - pz[j];
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best) { best = d2; arg = j; }
    }
    const float ga = sqrtf(best) - 2.f * c.phys.agent_radius;
    // ground: same support extent of the collision cylinder as the contact row of the step
    const float qx = b.state[3 * (size_t)S + s], qy = b.state[4 * (size_t)S + s];
    const float R22 = 1.f - 2.f * (qx * qx + qy * qy);
    const float ext = c.phys.col_radius * sqrtf(fmaxf(1.f - R22 * R22, 0.f)) + c.phys.col_halfheight * fabsf(R22) +
                      c.phys.col_margin;
    const float gg = z - ext - c.phys.ground_z;
    if (gap_agent) gap_agent[s] = ga;
    if (nearest) nearest[s] = arg;
    if (gap_ground) gap_ground[s] = gg;
    if (collision) collision[s] = (ga < thresh || gg < thresh) ? 1 : 0;
}

// rays: [R][3] directions in the body frame (body != 0) or world frame, start = pos + R * offset;
// hit_dist [S][R] (inf = no hit within range), hit_id [S][R]: -1 none, N = ground, j = agent j
__global__ void __launch_bounds__(128)
raycast_kernel(const __grid_constant__ MrsConfig c, const MrsBuffers b, const float* __restrict__ dirs, int nrays, float ox,
               float oy, float oz, int body, float range, float* __restrict__ hit_dist, int* __restrict__ hit_id) {
    const unsigned S = (unsigned)c.E * (unsigned)c.N;
    const unsigned tid = blockIdx.x * 128u + threadIdx.x;
    if (tid >= S * (unsigned)nrays) return;
    const unsigned s = tid / (unsigned)nrays;
    const int r = (int)(tid - s * (unsigned)nrays);
    const int N = c.N;
    const unsigned env0 = (s / (unsigned)N) * (unsigned)N;
    const int ai = (int)(s - env0);
    Agent st;
    load_agent(b.state, S, s, st);
    float Rm[9];
    quat_to_mat(st, Rm);
    float dx = dirs[3 * r], dy = dirs[3 * r + 1], dz = dirs[3 * r + 2];
    float sx = ox, sy = oy, sz = oz;
    if (body) {
        const float tx = Rm[0] * dx + Rm[1] * dy + Rm[2] * dz, ty = Rm[3] * dx + Rm[4] * dy + Rm[5] * dz,
                    tz = Rm[6] * dx + Rm[7] * dy + Rm[8] * dz;
        dx = tx; dy = ty; dz = tz;
        const float ux = Rm[0] * ox + Rm[1] * oy + Rm[2] * oz, uy = Rm[3] * ox + Rm[4] * oy + Rm[5] * oz,
                    uz = Rm[6] * ox + Rm[7] * oy + Rm[8] * oz;
        sx = ux; sy = uy; sz = uz;
    }
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
    dx *= inv; dy *= inv; dz *= inv;
    sx += st.px; sy += st.py; sz += st.pz;
    float best = range;
    int id = -1;
    // ground: top face of the 30 x 30 x 1 box centred at the origin (plane.urdf:21-26)
    if (dz < 0.f && sz > c.phys.ground_z) {
        const float t = (c.phys.ground_z - sz) / dz;
        const float hx = sx + t * dx, hy = sy + t * dy;
        if (t < best && fabsf(hx) <= 15.f && fabsf(hy) <= 15.f) { best = t; id = N; }
    }
    const float* px = b.state + env0;
    const float* py = b.state + (size_t)S + env0;
    const float* pz = b.state + 2 * (size_t)S + env0;
    const float rad2 = c.phys.agent_radius * c.phys.agent_radius;
    for (int j = 0; j < N; ++j) {
        if (j == ai) continue;
        const float cx = px[j] - sx, cy = py[j] - sy, cz = pz[j] - sz;
        const float tc = cx * dx + cy * dy + cz * dz;                  // closest approach along the ray
        const float d2 = cx * cx + cy * cy + cz * cz - tc * tc;
        if (d2 > rad2) continue;
        const float t = tc - sqrtf(rad2 - d2);
        if (t >= 0.f && t < best) { best = t; id = j; }
    }
    hit_dist[tid] = (id >= 0) ? best : INFINITY;
    hit_id[tid] = id;
}

// tape maintenance: 128-bit grid-stride copy / zero fill (slot sizes are multiples of 4 floats
// whenever E*N is; scalar tail otherwise)
__global__ void __launch_bounds__(256)
tape_fill_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t slot_elems, int count) {
    const size_t total = slot_elems * (size_t)count;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if ((slot_elems & 3) == 0) {
        const size_t n4 = total >> 2, s4 = slot_elems >> 2;
        float4* d4 = reinterpret_cast<float4*>(dst);
        const float4* r4 = reinterpret_cast<const float4*>(src);
        for (size_t i = tid; i < n4; i += stride) d4[i] = src ? r4[i % s4] : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (size_t i = tid; i < total; i += stride) dst[i] = src ? src[i % slot_elems] : 0.f;
    }
}

// ------------------------------------------------------------------------------ host side
static int g_sm_count = 0;

static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        g_sm_count = n;
    }
    return g_sm_count;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// largest float s with sqrt_rn(s) <= r  (see adjacency_pair)
static float adjacency_threshold(float r) {
    if (!(r >= 0.f)) return -1.f;          // negative or NaN range: nothing is adjacent
    if (isinf(r)) return INFINITY;
    float s = r * r;
    if (isinf(s)) return 3.402823466e+38f;  // every finite squared distance qualifies
    while (sqrtf(s) > r) s = nextafterf(s, -INFINITY);
    for (;;) {
        const float up = nextafterf(s, INFINITY);
        if (isinf(up) || sqrtf(up) > r) break;
        s = up;
    }
    return s;
}

static int check_cfg(const MrsConfig* cfg) {
    if (!cfg) return MRS_ERR_ARG;
    if (cfg->E <= 0 || cfg->N <= 0 || cfg->K < 0) return MRS_ERR_ARG;
    if (cfg->action_type < 0 || cfg->action_type > MRS_NO_ACTION) return MRS_ERR_ARG;
    if (cfg->state_layout < 0 || cfg->state_layout > MRS_X_FULL) return MRS_ERR_ARG;
    return MRS_OK;
}

static int last_error() { return cudaGetLastError() == cudaSuccess ? MRS_OK : MRS_ERR_CUDA; }

static Derived make_derived(const MrsConfig& c) {
    Derived d;
    const MrsQuadParams& q = c.quad;
    const MrsPhysicsParams& p = c.phys;
    d.inv_mass = (float)(1.0 / (double)p.mass);
    for (int i = 0; i < 3; ++i) d.inv_I[i] = (float)(1.0 / (double)p.inertia[i]);
    const double pr4 = (double)q.prop_radius / 4.0;
    d.gnd_c = (float)((double)q.kf * (double)q.gnd_eff_coeff * pr4 * pr4);
    d.dw_c = (float)((double)q.dw1 * pr4 * pr4);
    d.rpm2rad = (float)(2.0 * 3.14159265358979323846 / 60.0);
    d.q_x2 = (float)(0.25 * (double)c.dt * (double)c.dt);
    d.cap_w2 = (float)(((double)p.ang_motion_threshold / (double)c.dt) * ((double)p.ang_motion_threshold / (double)c.dt));
    const double cap_ang = 0.5 * 1.57079632679489661923 / (double)c.dt;      // Bullet: 0.5 * SIMD_HALF_PI / dt
    d.cap_k = (float)(sin(0.5 * cap_ang * (double)c.dt) / cap_ang);
    d.cap_c = (float)cos(0.5 * cap_ang * (double)c.dt);
    const float lim = 2.f * p.agent_radius + p.contact_margin;
    d.lim2 = lim * lim;
    d.gnd_skip_z = p.ground_z + p.contact_margin + p.col_radius + p.col_halfheight + p.col_margin + 1e-3f;
    d.inv_dt = (float)(1.0 / (double)c.dt);
    d.erp_dt = (float)((double)p.erp2 / (double)c.dt);
    d.comm_inf = isinf(c.comm_range) && c.comm_range > 0.f;
    d.s_max = adjacency_threshold(c.comm_range);
    d.inv_ctrl_dt = (float)(1.0 / (double)q.ctrl_dt);
    d.inv_4kf = (float)(1.0 / (4.0 * (double)q.kf));
    d.inv_pwm_a = (float)(1.0 / (double)q.pwm2rpm_a);
    d.inv_qmass = (float)(1.0 / (double)q.mass);
    return d;
}

// true iff the caller's configuration carries exactly the constants mrs_baked.cuh was generated from
static bool config_is_baked(const MrsConfig& c, const Derived& d) {
    MrsConfig cc = c;
    Derived dd = d;
    baked_constants(cc, dd);
    return memcmp(&cc, &c, sizeof(MrsConfig)) == 0 && memcmp(&dd, &d, sizeof(Derived)) == 0;
}

template <int MODE, int GT, int WPB, bool BAKED>
static int launch_group_wpb(const MrsConfig& c, const Derived& d, const MrsBuffers& b, const StepArgs& a, long long blocks,
                            bool pdl, cudaStream_t st) {
    constexpr size_t smem = (size_t)WPB * group_warp_smem_bytes<MODE, GT>();
    static bool configured[64] = {};          // per device: the attribute belongs to the function ON a device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return MRS_ERR_CUDA;
    if (!configured[dev]) {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(step_group_kernel<MODE, GT, WPB, BAKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
                cudaSuccess)
            return MRS_ERR_CUDA;
        configured[dev] = true;
    }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)blocks);
    lc.blockDim = dim3(WPB * 32);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl ? 1 : 0;
    if (cudaLaunchKernelEx(&lc, step_group_kernel<MODE, GT, WPB, BAKED>, c, d, b, a) != cudaSuccess) {
        (void)cudaGetLastError();
        return MRS_ERR_CUDA;
    }
    return last_error();
}

template <int MODE, int GT>
static int launch_group(const MrsConfig& c, const Derived& d, const MrsBuffers& b, const StepArgs& a, cudaStream_t st) {
    constexpr int kBig = 4 * ModeTraits<MODE>::minb;            // warps of a CTA that owns a whole SM
    const int sms = sm_count();
    if (sms <= 0) return MRS_ERR_CUDA;
    static const int use_pdl = env_int("MRS_B200_PDL", 1);
    static const int use_big = env_int("MRS_B200_BIGCTA", 1);
    // large jobs (every warp of the GPU gets more than two chunks): one SM-sized CTA per SM with the
    // shared-memory hand-out.  Programmatic dependent launch pays off for full waves (measured
    // -2.3 % at C5); partial waves are faster with plain stream order (C3: +11 % with PDL).
    static const int use_baked = env_int("MRS_B200_BAKED", 1);
    const bool baked = use_baked && config_is_baked(c, d);
    const int nwork = a.nchunks - a.chunk_lo;
    if (use_big && nwork > 2 * sms * kBig)
        return baked ? launch_group_wpb<MODE, GT, kBig, true>(c, d, b, a, sms, use_pdl != 0, st)
                     : launch_group_wpb<MODE, GT, kBig, false>(c, d, b, a, sms, use_pdl != 0, st);
    const long long need = ((long long)nwork + 3) / 4;
    const long long cap = (long long)sms * ModeTraits<MODE>::minb;
    const long long blocks = need < cap ? need : cap;
    const bool pdl = use_pdl && need >= cap;
    return baked ? launch_group_wpb<MODE, GT, 4, true>(c, d, b, a, blocks, pdl, st)
                 : launch_group_wpb<MODE, GT, 4, false>(c, d, b, a, blocks, pdl, st);
}

static int launch_adjacency(const float* pos, size_t cs, size_t as, float* A, int E, int N, float s_max, int comm_inf,
                            cudaStream_t st) {
    if (N >= 128 && (N & 3) == 0) {
        const int col_tiles = (N + 4 * kBlock - 1) / (4 * kBlock);
        const int row_tiles = (N + kRowTile - 1) / kRowTile;
        dim3 grid((unsigned)(col_tiles * row_tiles), (unsigned)E);
        adjacency_tiled_kernel<<<grid, kBlock, 0, st>>>(pos, cs, as, A, E, N, s_max, comm_inf);
    } else {
        const size_t total = (size_t)E * N * N;
        const size_t blocks = (total + 255) / 256;
        const size_t cap = (size_t)(sm_count() > 0 ? sm_count() : 148) * 32;
        adjacency_flat_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(pos, cs, as, A, E, N, s_max,
                                                                                        comm_inf);
    }
    return last_error();
}

// Side stream of the wide path: the adjacency kernel of step t (a pure streaming store that only
// reads the new positions) runs next to the compute-bound pair kernel of step t+1; it has to be done
// before post(t+1) overwrites the positions.  Fork / join through events, so it is capturable.
namespace {
struct SideLane {
    cudaStream_t s = nullptr;
    cudaEvent_t posted = nullptr, adj_done = nullptr;
    bool ok = false;
};
SideLane g_side[64];
SideLane* side_lane() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideLane& L = g_side[dev];
    if (!L.ok) {
        if (cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&L.posted, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&L.adj_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        L.ok = true;
    }
    return &L;
}
}  // namespace

// partner slices of pair_tile_kernel: ~8 CTAs per SM of a B200, at most kPairMaxSplit partial planes, whole
// tiles per slice.  A pure function of (E, N): mrs_scratch_planes sizes the caller's scratch from it.
static void pair_split(int E, int N, int* jw, int* nsplit) {
    const long long itiles = (N + kBlock - 1) / kBlock;
    const long long want = 8LL * 148;
    long long ns = (want + itiles * E - 1) / (itiles * E);
    if (ns > kPairMaxSplit) ns = kPairMaxSplit;
    if (ns > itiles) ns = itiles;                                                    // at least one tile per slice
    if (ns < 1) ns = 1;
    int w = (int)((N + ns - 1) / ns);
    w = (w + kBlock - 1) / kBlock * kBlock;
    *jw = w;
    *nsplit = (N + w - 1) / w;
}

template <int MODE, int LPA, int LPB>
static int launch_wide_lpa(const MrsConfig& c, const Derived& d, const MrsBuffers& b, const StepArgs& a, cudaStream_t st) {
    const size_t S = (size_t)c.E * c.N;
    constexpr int A = ModeTraits<MODE>::A;
    const unsigned blocks = (unsigned)((S * LPA + kBlock - 1) / kBlock);
    const unsigned blocks_post = (unsigned)((S * LPB + kBlock - 1) / kBlock);
    SideLane* L = (b.A_tape && a.T > 1) ? side_lane() : nullptr;
    int jw = 0, nsplit = 0;
    if (LPA > 32) pair_split(c.E, c.N, &jw, &nsplit);
    for (int t = 0; t < a.T; ++t) {
        const float* act_t = a.actions ? a.actions + (size_t)t * S * A : nullptr;
        if constexpr (LPA > 32) {
            const int itiles = (c.N + kBlock - 1) / kBlock;
            pair_tile_kernel<MODE><<<(unsigned)((long long)itiles * nsplit * c.E), kBlock, 0, st>>>(c, d, b, jw, nsplit);
            agent_pre_kernel<MODE><<<(unsigned)((S + kBlock - 1) / kBlock), kBlock, 0, st>>>(c, d, b, act_t, nsplit);
        } else {
            step_pre_kernel<MODE, LPA><<<blocks, kBlock, 0, st>>>(c, d, b, act_t);
        }
        if (L && t > 0 && cudaStreamWaitEvent(st, L->adj_done, 0) != cudaSuccess) return MRS_ERR_CUDA;
        step_post_kernel<LPB><<<blocks_post, kBlock, 0, st>>>(c, d, b, a.slot_x - t);
        if (b.A_tape) {
            cudaStream_t as = st;
            if (L) {
                if (cudaEventRecord(L->posted, st) != cudaSuccess) return MRS_ERR_CUDA;
                if (cudaStreamWaitEvent(L->s, L->posted, 0) != cudaSuccess) return MRS_ERR_CUDA;
                as = L->s;
            }
            const int rc = launch_adjacency(b.state, S, 1, b.A_tape + (size_t)(a.slot_a - t) * S * c.N, c.E, c.N, d.s_max,
                                            d.comm_inf, as);
            if (rc) return rc;
            if (L && cudaEventRecord(L->adj_done, L->s) != cudaSuccess) return MRS_ERR_CUDA;
        }
    }
    if (L && cudaStreamWaitEvent(st, L->adj_done, 0) != cudaSuccess) return MRS_ERR_CUDA;
    return last_error();
}

template <int MODE>
static int launch_tiled(const MrsConfig& c, const Derived& d, const MrsBuffers& b, const StepArgs& a, cudaStream_t st) {
    if (!b.scratch) return MRS_ERR_ARG;
    if ((unsigned long long)c.E * c.N * 32ull >= 0x7fffffffull * (unsigned long long)kBlock) return MRS_ERR_UNSUPPORTED;
    if (c.N <= 128) return launch_wide_lpa<MODE, 8, 8>(c, d, b, a, st);
    return launch_wide_lpa<MODE, 128, 32>(c, d, b, a, st);     // n-body tiles (pair_tile_kernel) + agent_pre_kernel
}

static int pow2ceil(int n) {
    int g = 1;
    while (g < n) g <<= 1;
    return g;
}

template <int MODE>
static int dispatch_step(const MrsConfig& c, const MrsBuffers& b, StepArgs a, cudaStream_t st) {
    const Derived d = make_derived(c);
    if (c.N <= 32) {
        a.G = pow2ceil(c.N);
        const size_t S = (size_t)c.E * c.N;
        a.xstride = (long long)(S * (size_t)state_dim(c.state_layout));
        a.astride = (long long)(S * (size_t)c.N);
        a.X0 = (b.X_tape && c.state_layout != MRS_X_NONE) ? b.X_tape + (size_t)a.slot_x * (size_t)a.xstride : nullptr;
        a.A0 = b.A_tape ? b.A_tape + (size_t)a.slot_a * (size_t)a.astride : nullptr;
        const int gpw = 32 / a.G;
        const int nchunks = (c.E + gpw - 1) / gpw;
        a.chunk_lo = 0;
        a.nchunks = nchunks;
        if (c.N != 8 && c.N != 16 && c.N != 32) return launch_group<MODE, 0>(c, d, b, a, st);
        // power-of-two swarms: unrolled pair loops over FULL chunks (32 valid slots); a ragged last chunk
        // (E not a multiple of 32/N) goes to the run-time-width kernel as a one-chunk launch
        const int nfull = c.E / gpw;
        int rc = MRS_OK;
        if (nfull > 0) {
            a.nchunks = nfull;
            switch (c.N) {
                case 8:  rc = launch_group<MODE, 8>(c, d, b, a, st); break;
                case 16: rc = launch_group<MODE, 16>(c, d, b, a, st); break;
                default: rc = launch_group<MODE, 32>(c, d, b, a, st); break;
            }
        }
        if (rc == MRS_OK && nfull < nchunks) {
            a.chunk_lo = nfull;
            a.nchunks = nchunks;
            rc = launch_group<MODE, 0>(c, d, b, a, st);
        }
        return rc;
    }
    return launch_tiled<MODE>(c, d, b, a, st);
}

static int step_impl(const MrsConfig* cfg, const MrsBuffers* bufs, const float* actions, int T, int slot_x, int slot_a,
                     void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bufs || !bufs->state || !bufs->ctrl) return MRS_ERR_ARG;
    if (T <= 0) return MRS_ERR_ARG;
    if (cfg->action_type != MRS_NO_ACTION && !actions) return MRS_ERR_ARG;
    if (bufs->X_tape && cfg->state_layout != MRS_X_NONE && (slot_x >= cfg->L || slot_x - (T - 1) < 0)) return MRS_ERR_ARG;
    if (bufs->A_tape && (slot_a >= cfg->L || slot_a - (T - 1) < 0)) return MRS_ERR_ARG;
    StepArgs a;
    a.actions = actions;
    a.T = T;
    a.slot_x = slot_x;
    a.slot_a = slot_a;
    a.G = 0;
    a.chunk_lo = 0;
    a.nchunks = 0;
    a.X0 = a.A0 = nullptr;
    a.xstride = a.astride = 0;
    if ((unsigned long long)cfg->E * cfg->N * (cfg->N > 18 ? cfg->N : 18) >= 0xffffffffull) return MRS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
#ifdef MRS_DEV_ONLY_MODE     // development builds: one action mode only (compile time / 8)
    if (cfg->action_type != MRS_DEV_ONLY_MODE) return MRS_ERR_UNSUPPORTED;
    return dispatch_step<MRS_DEV_ONLY_MODE>(*cfg, *bufs, a, st);
#else
    switch (cfg->action_type) {
        case MRS_SET_TARGET_VEL:   return dispatch_step<MRS_SET_TARGET_VEL>(*cfg, *bufs, a, st);
        case MRS_SET_TARGET_POS:   return dispatch_step<MRS_SET_TARGET_POS>(*cfg, *bufs, a, st);
        case MRS_SET_TARGET_ACCEL: return dispatch_step<MRS_SET_TARGET_ACCEL>(*cfg, *bufs, a, st);
        case MRS_SET_FORCE:        return dispatch_step<MRS_SET_FORCE>(*cfg, *bufs, a, st);
        case MRS_SET_TARGET_ORI:   return dispatch_step<MRS_SET_TARGET_ORI>(*cfg, *bufs, a, st);
        case MRS_SET_CONTROL:      return dispatch_step<MRS_SET_CONTROL>(*cfg, *bufs, a, st);
        case MRS_SET_SPEEDS:       return dispatch_step<MRS_SET_SPEEDS>(*cfg, *bufs, a, st);
        case MRS_NO_ACTION:        return dispatch_step<MRS_NO_ACTION>(*cfg, *bufs, a, st);
    }
    return MRS_ERR_ARG;
#endif
}

}  // namespace mrs

// ================================================================================ C ABI
using namespace mrs;

extern "C" {

int mrs_abi_version(void) { return MRS_ABI_VERSION; }

const char* mrs_strerror(int err) {
    switch (err) {
        case MRS_OK: return "ok";
        case MRS_ERR_ARG: return "invalid argument";
        case MRS_ERR_CUDA: return "CUDA error (no device, bad launch or asynchronous fault)";
        case MRS_ERR_UNSUPPORTED: return "unsupported configuration";
    }
    return "unknown error";
}

int mrs_action_dim(int action_type) {
    switch (action_type) {
        case MRS_SET_TARGET_VEL: case MRS_SET_TARGET_POS: case MRS_SET_TARGET_ACCEL: case MRS_SET_FORCE:
        case MRS_SET_TARGET_ORI: return 3;
        case MRS_SET_CONTROL: case MRS_SET_SPEEDS: return 4;
    }
    return 0;
}

int mrs_state_dim(int state_layout) { return state_dim(state_layout); }

int mrs_scratch_planes(int E, int N) {
    if (N <= 32 || E <= 0) return 0;
    if (N <= 128) return MRS_SCRATCH_PLANES;
    int jw = 0, nsplit = 0;
    pair_split(E, N, &jw, &nsplit);
    return MRS_SCRATCH_PLANES + 2 * nsplit;
}

size_t mrs_sizeof_config(void) { return sizeof(MrsConfig); }
size_t mrs_sizeof_buffers(void) { return sizeof(MrsBuffers); }

int mrs_default_config(MrsConfig* cfg) {
    if (!cfg) return MRS_ERR_ARG;
    memset(cfg, 0, sizeof(*cfg));
    cfg->E = 1; cfg->N = 1; cfg->K = 0; cfg->L = 1;
    cfg->action_type = MRS_SET_TARGET_VEL;
    cfg->state_layout = MRS_X_POS_VEL;
    cfg->dt = 0.01f;
    cfg->gravity = 9.81f;
    cfg->comm_range = INFINITY;
    MrsQuadParams& q = cfg->quad;
    // cf2x.urdf:5,11-12 and prop link CoM offsets :42,54,66,78
    q.mass = 0.027f; q.ixx = 1.4e-5f; q.iyy = 1.4e-5f; q.izz = 2.17e-5f;
    q.kf = 3.16e-10f; q.km = 7.94e-12f; q.arm = 0.0397f;
    q.gnd_eff_coeff = 11.36859f; q.prop_radius = 2.31348e-2f;
    q.drag_xy = 9.1785e-7f; q.drag_z = 10.311e-7f;
    q.dw1 = 2267.18f; q.dw2 = 0.16f; q.dw3 = -0.11f;
    const float px[4] = {0.028f, -0.028f, -0.028f, 0.028f}, py[4] = {0.028f, 0.028f, -0.028f, -0.028f};
    for (int i = 0; i < 4; ++i) { q.prop_x[i] = px[i]; q.prop_y[i] = py[i]; }
    // Quadcopter.calculate_parameters (Quadcopter.py:153-168), in double then rounded
    {
        const double g = 9.81 * 0.027, kf = 3.16e-10, t2w = 2.25, coeff = 11.36859, pr = 2.31348e-2;
        const double max_rpm = sqrt(t2w * g / (4 * kf));
        const double max_thrust = 4.0 * kf * max_rpm * max_rpm;
        q.gnd_hclip = (float)(0.25 * pr * sqrt((15.0 * max_rpm * max_rpm * kf * coeff) / max_thrust));
    }
    // QuadControl gains (QuadControl.py:14-32)
    q.pos_p = 1.5f; q.pos_i = 0.001f; q.pos_d = 1.0f;
    q.vel_p = 3.0f; q.vel_i = 0.1f; q.vel_d = 1.0f;
    const float op[3] = {70000.f, 70000.f, 60000.f}, oi[3] = {0.f, 0.f, 500.f}, od[3] = {20000.f, 20000.f, 12000.f};
    for (int i = 0; i < 3; ++i) { q.ori_p[i] = op[i]; q.ori_i[i] = oi[i]; q.ori_d[i] = od[i]; }
    q.min_pwm = 20000.f; q.max_pwm = 65535.f; q.pwm2rpm_a = 0.2685f; q.pwm2rpm_b = 4070.3f;
    q.ctrl_dt = 0.01f; q.ctrl_gravity = 9.81f;
    // 'x' mixer (Quadcopter.py:164) and its inverse; orthogonal rows => Ainv = A^T D^-1
    const double r2 = 1.0 / sqrt(2.0);
    const double A[4][4] = {{1, 1, 1, 1}, {r2, r2, -r2, -r2}, {-r2, r2, r2, -r2}, {-1, 1, -1, 1}};
    const double rown[4] = {4.0, 2.0, 2.0, 4.0};
    for (int r = 0; r < 4; ++r)
        for (int cidx = 0; cidx < 4; ++cidx) {
            q.mix_a[r * 4 + cidx] = (float)A[r][cidx];
            q.mix_ainv[cidx * 4 + r] = (float)(A[r][cidx] / rown[r]);
        }
    // least-squares solve matrices of the 16 active sets: P_S = (A_S^T A_S)^-1 A_S^T (4x4, zero rows off S)
    for (int m = 0; m < 16; ++m) {
        int cols[4], nc = 0;
        for (int cidx = 0; cidx < 4; ++cidx) if ((m >> cidx) & 1) cols[nc++] = cidx;
        double Gm[4][8];
        for (int i = 0; i < nc; ++i) {
            for (int j = 0; j < nc; ++j) {
                double acc = 0;
                for (int r = 0; r < 4; ++r) acc += A[r][cols[i]] * A[r][cols[j]];
                Gm[i][j] = acc;
            }
            for (int j = 0; j < nc; ++j) Gm[i][nc + j] = (i == j) ? 1.0 : 0.0;
        }
        for (int p = 0; p < nc; ++p) {   // Gauss-Jordan, SPD Gram matrix
            int best = p;
            for (int r = p + 1; r < nc; ++r) if (fabs(Gm[r][p]) > fabs(Gm[best][p])) best = r;
            if (best != p) for (int j = 0; j < 2 * nc; ++j) { double tmp = Gm[p][j]; Gm[p][j] = Gm[best][j]; Gm[best][j] = tmp; }
            const double piv = Gm[p][p];
            for (int j = 0; j < 2 * nc; ++j) Gm[p][j] /= piv;
            for (int r = 0; r < nc; ++r) if (r != p) {
                const double f = Gm[r][p];
                for (int j = 0; j < 2 * nc; ++j) Gm[r][j] -= f * Gm[p][j];
            }
        }
        for (int i = 0; i < nc; ++i)
            for (int r = 0; r < 4; ++r) {
                double acc = 0;
                for (int j = 0; j < nc; ++j) acc += Gm[i][nc + j] * A[r][cols[j]];
                q.nnls_tab[m * 16 + cols[i] * 4 + r] = (float)acc;
            }
    }
    MrsPhysicsParams& p = cfg->phys;   // oracle/bullet_model.py PhysicsParams
    p.mass = 0.027f;
    {
        const double hx = 0.06 + 3 * 0.001, hz = 0.0125 + 3 * 0.001, lx = 2 * hx, lz = 2 * hz, m = 0.027;
        p.inertia[0] = p.inertia[1] = (float)(m / 12.0 * (lx * lx + lz * lz));
        p.inertia[2] = (float)(m / 12.0 * (lx * lx + lx * lx));
    }
    p.lin_damping = 0.04f; p.ang_damping = 0.04f; p.max_coord_vel = 100.f; p.gyro = 1;
    p.ang_motion_threshold = 0.78539816339744830962f;
    p.erp2 = 0.08f; p.slop = 1e-5f; p.contact_margin = 0.02f;
    p.mu_ground = 0.75f; p.ground_z = 0.5f;
    p.col_radius = 0.06f; p.col_halfheight = 0.0125f; p.col_margin = 0.001f;
    p.ground_contact = 1; p.agent_contact = 1;
    p.agent_radius = 0.3f;
    return MRS_OK;
}

int mrs_config_is_baked(const MrsConfig* cfg) {
    if (check_cfg(cfg)) return 0;
    return config_is_baked(*cfg, make_derived(*cfg)) ? 1 : 0;
}

int mrs_debug_derived(const MrsConfig* cfg, void* out, size_t out_bytes) {
    if (!cfg || !out || out_bytes != sizeof(Derived)) return MRS_ERR_ARG;
    const Derived d = make_derived(*cfg);
    memcpy(out, &d, sizeof(Derived));
    return MRS_OK;
}

int mrs_step(const MrsConfig* cfg, const MrsBuffers* bufs, const float* actions, int slot_x, int slot_a, void* stream) {
    return step_impl(cfg, bufs, actions, 1, slot_x, slot_a, stream);
}

int mrs_step_many(const MrsConfig* cfg, const MrsBuffers* bufs, const float* actions, int T, int slot_x_first,
                  int slot_a_first, void* stream) {
    return step_impl(cfg, bufs, actions, T, slot_x_first, slot_a_first, stream);
}

int mrs_observe(const MrsConfig* cfg, const MrsBuffers* bufs, int slot, int write_X, int write_A, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bufs || !bufs->state) return MRS_ERR_ARG;
    if (slot < 0 || slot >= cfg->L) return MRS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t S = (size_t)cfg->E * cfg->N;
    if (write_X && cfg->state_layout != MRS_X_NONE) {
        if (!bufs->X_tape) return MRS_ERR_ARG;
        observe_x_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(*bufs, S, cfg->state_layout, slot);
        if ((rc = last_error())) return rc;
    }
    if (write_A) {
        if (!bufs->A_tape) return MRS_ERR_ARG;
        const int comm_inf = isinf(cfg->comm_range) && cfg->comm_range > 0.f;
        rc = launch_adjacency(bufs->state, S, 1, bufs->A_tape + (size_t)slot * S * cfg->N, cfg->E, cfg->N,
                              adjacency_threshold(cfg->comm_range), comm_inf, st);
    }
    return rc;
}

int mrs_adjacency(const MrsConfig* cfg, const float* pos, float* A, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!pos || !A) return MRS_ERR_ARG;
    const int comm_inf = isinf(cfg->comm_range) && cfg->comm_range > 0.f;
    return launch_adjacency(pos, 1, 3, A, cfg->E, cfg->N, adjacency_threshold(cfg->comm_range), comm_inf,
                            (cudaStream_t)stream);
}

int mrs_set_state(const MrsConfig* cfg, const MrsBuffers* bufs, const float* pos, const float* ori_euler,
                  const float* vel, const float* angvel, const unsigned char* env_mask, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bufs || !bufs->state) return MRS_ERR_ARG;
    const size_t S = (size_t)cfg->E * cfg->N;
    set_state_kernel<<<(unsigned)((S + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bufs->state, S, cfg->N, pos,
                                                                                   ori_euler, vel, angvel, env_mask);
    return last_error();
}

int mrs_tape_fill(const MrsConfig* cfg, const MrsBuffers* bufs, int which, int src, int dst_first, int count,
                  void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bufs || (which != 1 && which != 2)) return MRS_ERR_ARG;
    if (count <= 0) return count == 0 ? MRS_OK : MRS_ERR_ARG;
    if (dst_first < 0 || dst_first + count > cfg->L || src >= cfg->L) return MRS_ERR_ARG;
    if (src >= 0 && src >= dst_first && src < dst_first + count) return MRS_ERR_ARG;
    const size_t S = (size_t)cfg->E * cfg->N;
    float* tape = (which == 1) ? bufs->X_tape : bufs->A_tape;
    if (!tape) return MRS_ERR_ARG;
    const size_t slot_elems = (which == 1) ? S * (size_t)state_dim(cfg->state_layout) : S * (size_t)cfg->N;
    if (slot_elems == 0) return MRS_ERR_ARG;
    const size_t work = (slot_elems * count + 3) / 4;
    size_t blocks = (work + 255) / 256;
    const size_t cap = (size_t)(sm_count() > 0 ? sm_count() : 148) * 16;
    if (blocks > cap) blocks = cap;
    tape_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        tape + (size_t)dst_first * slot_elems, src >= 0 ? tape + (size_t)src * slot_elems : nullptr, slot_elems, count);
    return last_error();
}

int mrs_step_host(const MrsConfig* cfg, const MrsBuffers* bufs, const float* actions_host, float* dev_actions,
                  float* X_host, float* A_host, int slot_x, int slot_a, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bufs) return MRS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t S = (size_t)cfg->E * cfg->N;
    const int adim = mrs_action_dim(cfg->action_type);
    if (adim > 0) {
        if (!actions_host || !dev_actions) return MRS_ERR_ARG;
        if (cudaMemcpyAsync(dev_actions, actions_host, S * adim * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess)
            return MRS_ERR_CUDA;
    }
    rc = step_impl(cfg, bufs, dev_actions, 1, slot_x, slot_a, stream);
    if (rc) return rc;
    const int D = state_dim(cfg->state_layout);
    if (X_host && bufs->X_tape && D > 0) {
        if (cudaMemcpyAsync(X_host, bufs->X_tape + (size_t)slot_x * S * D, S * D * sizeof(float), cudaMemcpyDeviceToHost,
                            st) != cudaSuccess)
            return MRS_ERR_CUDA;
    }
    if (A_host && bufs->A_tape) {
        if (cudaMemcpyAsync(A_host, bufs->A_tape + (size_t)slot_a * S * cfg->N, S * cfg->N * sizeof(float),
                            cudaMemcpyDeviceToHost, st) != cudaSuccess)
            return MRS_ERR_CUDA;
    }
    return cudaStreamSynchronize(st) == cudaSuccess ? MRS_OK : MRS_ERR_CUDA;
}


// Pipelined host rollout: three streams so that the H2D of step t+1's actions and the D2H of step
// t-1's X / A slices overlap the kernel of step t (PCIe is full duplex).  The compute stream is the
// caller's; the two copy streams and four events are created once per device and reused.
namespace {
struct CopyLanes {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, tail = nullptr;
    bool ok = false;
};
CopyLanes g_lanes[64];
CopyLanes* copy_lanes() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    CopyLanes& L = g_lanes[dev];
    if (!L.ok) {
        if (cudaStreamCreateWithFlags(&L.h2d, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&L.d2h, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 2; ++i) {
            if (cudaEventCreateWithFlags(&L.up[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&L.done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&L.tail, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        L.ok = true;
    }
    return &L;
}
}  // namespace

int mrs_rollout_host(const MrsConfig* cfg, const MrsBuffers* bufs, const float* actions_host, float* dev_actions,
                     float* X_host, float* A_host, int T, int slot_x_first, int slot_a_first, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bufs || T <= 0) return MRS_ERR_ARG;
    const int adim = mrs_action_dim(cfg->action_type);
    if (adim <= 0 || !actions_host || !dev_actions) return MRS_ERR_ARG;
    CopyLanes* L = copy_lanes();
    if (!L) return MRS_ERR_CUDA;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t S = (size_t)cfg->E * cfg->N;
    const size_t abytes = S * adim * sizeof(float);
    const int D = mrs_state_dim(cfg->state_layout);
    const size_t xelems = S * (size_t)D, aelems = S * (size_t)cfg->N;
    // the copy lanes start after whatever the caller already queued on the compute stream
    if (cudaEventRecord(L->tail, st) != cudaSuccess) return MRS_ERR_CUDA;
    if (cudaStreamWaitEvent(L->h2d, L->tail, 0) != cudaSuccess) return MRS_ERR_CUDA;
    if (cudaStreamWaitEvent(L->d2h, L->tail, 0) != cudaSuccess) return MRS_ERR_CUDA;
    for (int t = 0; t < T; ++t) {
        const int bsel = t & 1;
        float* da = dev_actions + (size_t)bsel * S * adim;
        if (t >= 2 && cudaStreamWaitEvent(L->h2d, L->done[bsel], 0) != cudaSuccess) return MRS_ERR_CUDA;
        if (cudaMemcpyAsync(da, actions_host + (size_t)t * S * adim, abytes, cudaMemcpyHostToDevice, L->h2d) != cudaSuccess)
            return MRS_ERR_CUDA;
        if (cudaEventRecord(L->up[bsel], L->h2d) != cudaSuccess) return MRS_ERR_CUDA;
        if (cudaStreamWaitEvent(st, L->up[bsel], 0) != cudaSuccess) return MRS_ERR_CUDA;
        rc = step_impl(cfg, bufs, da, 1, slot_x_first - t, slot_a_first - t, stream);
        if (rc) return rc;
        if (cudaEventRecord(L->done[bsel], st) != cudaSuccess) return MRS_ERR_CUDA;
        if ((X_host && bufs->X_tape && D > 0) || (A_host && bufs->A_tape)) {
            if (cudaStreamWaitEvent(L->d2h, L->done[bsel], 0) != cudaSuccess) return MRS_ERR_CUDA;
            if (X_host && bufs->X_tape && D > 0 &&
                cudaMemcpyAsync(X_host + (size_t)t * xelems, bufs->X_tape + (size_t)(slot_x_first - t) * xelems,
                                xelems * sizeof(float), cudaMemcpyDeviceToHost, L->d2h) != cudaSuccess)
                return MRS_ERR_CUDA;
            if (A_host && bufs->A_tape &&
                cudaMemcpyAsync(A_host + (size_t)t * aelems, bufs->A_tape + (size_t)(slot_a_first - t) * aelems,
                                aelems * sizeof(float), cudaMemcpyDeviceToHost, L->d2h) != cudaSuccess)
                return MRS_ERR_CUDA;
        }
    }
    // join: the caller's stream continues only after the last copies
    if (cudaEventRecord(L->tail, L->d2h) != cudaSuccess) return MRS_ERR_CUDA;
    if (cudaStreamWaitEvent(st, L->tail, 0) != cudaSuccess) return MRS_ERR_CUDA;
    return cudaStreamSynchronize(st) == cudaSuccess ? MRS_OK : MRS_ERR_CUDA;
}


int mrs_spawn(const MrsConfig* cfg, const MrsBuffers* bufs, const unsigned char* env_mask, unsigned long long seed,
              float z_lo, float z_hi, float xy_radius, float xy_sigma, float yaw_lo, float yaw_hi, int max_rounds,
              unsigned int* failed_envs, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bufs || !bufs->state) return MRS_ERR_ARG;
    if (cfg->N > 32) return MRS_ERR_UNSUPPORTED;
    if (!(z_hi >= z_lo) || !(xy_radius > 0.f) || !(xy_sigma > 0.f) || max_rounds <= 0) return MRS_ERR_ARG;
    SpawnArgs sp;
    sp.seed = seed; sp.z_lo = z_lo; sp.z_hi = z_hi; sp.xy_radius = xy_radius; sp.xy_sigma = xy_sigma;
    sp.yaw_lo = yaw_lo; sp.yaw_hi = yaw_hi; sp.max_rounds = max_rounds;
    spawn_kernel<<<(unsigned)((cfg->E + 3) / 4), 128, 0, (cudaStream_t)stream>>>(*cfg, *bufs, sp, env_mask, failed_envs);
    return last_error();
}


int mrs_proximity(const MrsConfig* cfg, const MrsBuffers* bufs, float threshold, float* gap_agent, int* nearest,
                  float* gap_ground, unsigned char* collision, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bufs || !bufs->state) return MRS_ERR_ARG;
    const size_t S = (size_t)cfg->E * cfg->N;
    proximity_kernel<<<(unsigned)((S + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*cfg, *bufs, threshold, gap_agent,
                                                                                   nearest, gap_ground, collision);
    return last_error();
}

int mrs_raycast(const MrsConfig* cfg, const MrsBuffers* bufs, const float* directions, int n_rays, const float* offset3,
                int body_frame, float range, float* hit_dist, int* hit_id, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bufs || !bufs->state || !directions || n_rays <= 0 || !hit_dist || !hit_id || !(range > 0.f)) return MRS_ERR_ARG;
    const size_t total = (size_t)cfg->E * cfg->N * n_rays;
    if (total >= 0xffffffffull) return MRS_ERR_UNSUPPORTED;
    const float ox = offset3 ? offset3[0] : 0.f, oy = offset3 ? offset3[1] : 0.f, oz = offset3 ? offset3[2] : 0.f;
    raycast_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*cfg, *bufs, directions, n_rays, ox, oy,
                                                                                     oz, body_frame, range, hit_dist, hit_id);
    return last_error();
}

}  // extern "C"
