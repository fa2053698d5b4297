// sim_kernels.cuh -- sm_100a kernels of the lock-step traffic microsimulation.
//
// One CTA owns one environment instance for a whole MultiSignal.step() (multi_signal.py:164-197):
// the instance's vehicle tile (Structure-of-Arrays, lane-major, front vehicle first) is staged
// from HBM into shared memory once, `step_length` one-second ticks run entirely out of shared
// memory (car-following / lane change / junction right-of-way / lane hand-off / insertion /
// arrival), Signal.observe + states.* + rewards.* are reduced with warp shuffles, and the tile is
// written back.  HBM traffic per env step is therefore ~1 read + 1 write of the tile instead of
// one per tick.  No tensor cores: there is no dense contraction on this path.
//
// Arithmetic is plain IEEE fp32 with FMA contraction disabled (-fmad=false), the same operation
// order as the rule set written down in DESIGN.md §4, so results are bit-identical to the CPU
// oracle used by the tests.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/resco_b200.h"

#ifndef RS_INLINE_HEAVY
#define RS_INLINE_HEAVY 1
#endif
#if RS_INLINE_HEAVY
#define RS_HEAVY __device__ __forceinline__
#else
#define RS_HEAVY __device__ __noinline__
#endif

namespace rs {

constexpr int kMaxHops = 8;
constexpr float kHaltSpeed = 0.1f;
constexpr float kNumEps = 0.001f;
constexpr float kEmergencyDecel = 9.0f;
constexpr int kLcCooldown = 5;
constexpr float kCoopMargin = 1.0f;   // extra room a yielding follower leaves behind an urgent lane changer (oracle: COOP_MARGIN)
constexpr int kVehWords = 11;      // 32-bit words per vehicle record
constexpr int kHdrInts = 16;
constexpr uint32_t kArrived = 0xFFFFu;

// header slots (ints; floats bit-cast)
// H_DONE: launch stamp (launch epoch + 1) of the last launch that stepped the instance; it sits in the same 32-byte
// sector as H_NVEH, so a CTA staging the header sees both from the same write-back
enum { H_TICK = 0, H_NVEH, H_EPOCH, H_NINS, H_NARR, H_ANOM, H_ACTIVE, H_DONE, H_F_DELAY_ARR, H_F_DUR_ARR, H_F_PENDING, H_NREF, H_F_WAIT_ARR };

// vtype table columns
enum { VT_LEN = 0, VT_GAP, VT_ACCEL, VT_DECEL, VT_TAU, VT_SIGMA, VT_VMAX, VT_DEV };

// Device-side network tables.  rs_create packs the flat per-field arrays of RsScenario (the ABI) into one record per
// lane / link / foe, so that everything the junction logic needs about one object sits in one or two 32-byte sectors
// (one L1 miss instead of a dozen) and the dependent load chains of the right-of-way loops are one level deep.
struct alignas(16) LaneRec {
  float len, vmax; int32_t link_off, internal;
  int32_t index, left, right, perm;
  float tls_dist; int32_t sig, sig_slot, pad;
};                                                   // 48 B; [n_lanes + 1] (sentinel carries link_off = n_links)
struct alignas(16) LinkRec {
  int32_t from, to, via, to_edge;
  int32_t tls, tlidx, state, cont;
  float via_len; int32_t last_int, parent, foe_off;
  int32_t nxt /* via >= 0 ? via : to */;
  // the last internal lanes of ALL foes of this link as a bit mask over two consecutive words of the per-instance
  // lane-occupancy bit array (Tile::occ), so "is somebody crossing my path" is two ANDs instead of a loop over
  // the foes; occ_word = -1 if the foes' lanes do not fit such a window (then the loop runs)
  int32_t occ_word; uint32_t occ_lo, occ_hi;
  // what a hop of the junction look-ahead needs about the lanes on both sides of the link, joined in, so that a hop
  // touches this record instead of three or four (on the big maps every table line is an L2 round trip: two 110 KB CTAs
  // leave ~30 KB of L1 for 0.8 MB of tables)
  float nxt_len, nxt_vmax;                           // lane_len / lane_vmax of nxt
  int32_t nxt_internal, from_internal;               // lane_internal of nxt / of from
  int32_t nxt_link;                                  // nxt internal: its one continuing link, or -2 if it has none; nxt normal: -3 (route lookup)
  int32_t yield_parent;                              // from internal: the entry link whose foes are yielded to at this point
                                                     // (parent with an internal junction whose waiting slot is `from`), else -1
  float yield_cross;                                 // ... and its crossing distance without the vehicle length
  int32_t foe_end;                                   // foe_off of the next link
};                                                   // 96 B; [n_links + 1] (sentinel carries foe_off = n_foes)
struct alignas(16) FoeRec {                          // one foe of a link, joined with what link_blocked reads of it
  int32_t link, flags, last_int, from;               // foe link f, foe_flags, link_last_int[f], link_from[f]
  int32_t slot /* link_cont[f] ? link_via[f] : -1 */; float via_len, len_from, len_slot;
};                                                   // 32 B; [n_foes + 1] (one readable record past the end)

struct DevScenario : RsScenario {                    // base-class pointers are DEVICE pointers
  const LaneRec* lane_rec;
  const LinkRec* link_rec;
  const FoeRec* foe_rec;
  // choose_link() tabulated: [n_route_steps][8] = encode_nextlink code of the link a vehicle at route step s takes
  // from lane index j of that step's edge (0xFE route ends, 0xFD the lane does not lead on)
  const uint8_t* route_step_link;
  int32_t tile_cap;      // vehicles the tile of THIS launch holds (<= vcap; vcap stays the stride of the HBM store).  A
                         // launch whose tile is smaller than the store defers an instance that outgrows it to the
                         // overflow pass (DevSim::overflow_*), which runs it again from the untouched HBM state
  int32_t tile_single;   // one tile buffer in shared memory (see SmemLayout::single)
  int32_t redo_cap;      // > 0 (launches with several instances per CTA): an instance that outgrows its slot's tile is
                         // stepped again at once by the WHOLE CTA on a tile of redo_cap vehicles laid over the CTA's
                         // shared memory (all G x TPI threads on one instance); only what outgrows that too is deferred
  int32_t redo_single;   // the redo tile is single-buffered
  int32_t tile_gmem;     // the vehicle tile and the per-vehicle scratch live in a per-CTA global-memory workspace (L2
                         // resident) instead of shared memory: vehicle stores larger than one CTA's shared memory
};

// Device copy of the scenario + per-sim buffers.
struct DevSim {
  DevScenario sc;
  int32_t n_env;
  uint64_t seed;
  int64_t first_env_id;
  // per-instance state in HBM
  int32_t* hdr;           // [N][kHdrInts]
  int32_t* tls_phase;     // [N][n_tls]
  int32_t* tls_end;       // [N][n_tls]
  int32_t* next_phase;    // [N][S]
  int32_t* origin_cur;    // [N][O]
  int32_t* origin_backlog;// [N][O]
  uint32_t* veh;          // [N][kVehWords][vcap]
  // observation outputs
  float *lane_queue, *lane_approach, *lane_total_wait, *lane_max_wait, *lane_speed_sum, *lane_arrivals;  // [N][SL]
  int32_t* phase_obs;     // [N][S]
  float *mplight, *wave, *rew_wait, *rew_wait_norm, *rew_pressure;
  int32_t *sig_queue_len, *sig_max_queue;
  int4* trip_rec;         // [N][n_trips] {arrival tick, depart tick, timeLoss bits, depart delay} or null
  float* trip_wait;       // [N][n_trips] tripinfo waitingTime (with trip_rec)
  float *drq, *drq_norm;  // [N][SL][5] optional (out_mask)
  float* mplight_full;    // [N][S][49] optional
  int32_t out_mask;       // RS_OUT_* bits
  int32_t* work_counter;  // dynamic instance scheduler of the persistent launch
  unsigned char* workspace;   // tile_gmem: [grid CTAs x instances per CTA][SmemLayout::veh_total] bytes
  int32_t* overflow_count;    // instances the fast pass deferred (tile outgrown); null: this launch does not defer
  int32_t* overflow_list;     // [N] their local ids
  int32_t* redo_count;        // instances stepped again inside their CTA (redo_cap) during this launch
  // Heavy list (launches with an in-CTA redo tile): instances that END a launch with more than tile_cap - heavy_margin
  // vehicles are listed for the next launch, which steps them FIRST, each by a whole CTA on the redo tile, while the
  // other CTAs start on the groups -- the long items lead, the short ones fill in (no tail behind a late redo).
  int32_t* heavy_list[2];     // [N] each; [heavy_cur] is read by this launch, the other one is filled by it
  int32_t* heavy_count;       // [4]: [0], [1] entries of the two lists, [2] = heavy_cur, the list THIS launch reads, [3] launch
                              // epoch (device side so that a replayed CUDA graph sees them; k_heavy_flip after every launch)
  int32_t* heavy_taken;       // work counter over heavy_list[heavy_cur]
  int32_t heavy_thr;          // fast tile - margin: the same threshold in every pass / run of the launch
  int32_t from_list;          // this launch IS the overflow pass: instance ids come from overflow_list[0 .. *overflow_count)
  unsigned long long* phase_clocks;   // [24] diagnostics (RS_PHASE_CLOCKS builds)
  int32_t persistent;
  int32_t use_tma;        // stage the tile with cp.async.bulk (TMA 1-D bulk copies) instead of LDG/STG
};

struct RunArgs {
  const int32_t* actions;  // [N][S] or null
  int32_t do_prep;         // Signal.prep_phase from actions
  int32_t ticks_a;         // ticks before set_phase
  int32_t do_set;          // Signal.set_phase(next_phase)
  int32_t ticks_b;         // ticks after
  int32_t do_observe;
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11)
__device__ __forceinline__ void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r > 0) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  }
}
enum { STREAM_SF = 1, STREAM_DAWDLE = 2, STREAM_DEMAND = 3, STREAM_ROUTE = 4, STREAM_DEPARTPOS = 5 };

// ---------------------------------------------------------------------------------------------
// car-following primitives (SUMO Euler update, dt = 1 s) -- see DESIGN.md §4.2
__device__ __forceinline__ float brake_gap(float speed, float decel, float headway) {
  int steps = (int)(speed / decel);
  float fs = (float)steps;
  float a = fs * speed;
  float b = decel * fs;
  float c = b * (float)(steps + 1);
  float d = c / 2.0f;
  return (a - d) + speed * headway;
}
__device__ __forceinline__ float max_safe_stop_speed(float gap, float decel, float tau) {
  float g = gap - kNumEps;
  if (g < 0.0f) return 0.0f;
  float b = decel, t = tau;
  float q = 2.0f * g / b - t;
  float inner = 1.0f + 4.0f * (q + t * t);
  float n = floorf(0.5f - ((t + sqrtf(inner) * -0.5f)));
  float h = 0.5f * n * (n - 1.0f) * b + n * b * t;
  float r = (g - h) / (n + t);
  return n * b + r;
}
__device__ __forceinline__ float follow_speed(float gap, float vlead, float dlead, float decel, float tau) {
  float bg = brake_gap(vlead, fmaxf(decel, dlead), 0.0f);
  return max_safe_stop_speed(gap + bg, decel, tau);
}
__device__ __forceinline__ float free_speed(float decel, float dist, float target) {
  if (dist < target) return target;
  float b = decel;
  float bb = b + 2.0f * target;
  float y = fmaxf(0.0f, ((sqrtf(bb * bb + 8.0f * b * dist) - b) * 0.5f - target) / b);
  float yf = floorf(y);
  float exact = (yf * yf * b + yf * b) / 2.0f + yf * target + (y > yf ? target : 0.0f);
  return fmaxf(0.0f, dist - exact) / (yf + 1.0f) + yf * b + target;
}
__device__ __forceinline__ float dawdle(float v, float accel, float sigma, float xi) {
  if (v < accel) v -= sigma * v * xi; else v -= sigma * accel * xi;
  return fmaxf(0.0f, v);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory view of one instance
struct Tile {
  // vehicle SoA (current buffer)
  float* pos; float* speed; float* sf; float* tloss;
  int32_t* vid;
  uint32_t* wr;    // wait (lo16) | rwait (hi16), whole seconds
  uint32_t* rc;    // route (lo16) | cursor (hi16)
  uint32_t* meta;  // vtype (8) | lcc (8) | seen_sig (8, 0xFF none) | spare
  uint32_t* ed;    // seen_epoch (lo16) | depart tick (hi16)
  uint32_t* dl;    // depart delay (lo16) | lane (hi16)
  uint32_t* aw;    // accumulated waiting seconds, the tripinfo waitingTime (lo16) | spare (hi16)
  uint16_t* lane_start;   // [L+1]
  const uint32_t* occ;    // [(L+31)/32 + 2] bit l = lane l holds a vehicle (state at the start of the tick)
  int32_t* tls_phase; int32_t* tls_end;
  int32_t* tls_state;     // [n_tls] offset into state_chars of the phase currently shown
  const float* vt;        // vtype table in smem
  int32_t tick;
  uint32_t env_lo, env_hi, seed_lo, seed_hi;
};

__device__ __forceinline__ void tile_bind(Tile& t, uint32_t* base, int vcap) {
  t.pos = (float*)(base + 0 * vcap); t.speed = (float*)(base + 1 * vcap); t.sf = (float*)(base + 2 * vcap);
  t.tloss = (float*)(base + 3 * vcap); t.vid = (int32_t*)(base + 4 * vcap); t.wr = base + 5 * vcap;
  t.rc = base + 6 * vcap; t.meta = base + 7 * vcap; t.ed = base + 8 * vcap; t.dl = base + 9 * vcap;
  t.aw = base + 10 * vcap;
}

#define VTT(t, i, f) ((t).vt[(i) * 8 + (f)])
__device__ __forceinline__ int v_vtype(const Tile& t, int i) { return (int)(t.meta[i] & 0xFFu); }
__device__ __forceinline__ int v_lcc(const Tile& t, int i) { return (int)((t.meta[i] >> 8) & 0xFFu); }
__device__ __forceinline__ int v_route(const Tile& t, int i) { return (int)(t.rc[i] & 0xFFFFu); }
__device__ __forceinline__ int v_cursor(const Tile& t, int i) { return (int)(t.rc[i] >> 16); }
__device__ __forceinline__ int v_wait(const Tile& t, int i) { return (int)(t.wr[i] & 0xFFFFu); }
__device__ __forceinline__ int v_lane(const Tile& t, int i) { return (int)(t.dl[i] >> 16); }
// cached choose_link() of the vehicle's CURRENT lane (meta bits 24..31: index relative to the lane's
// first link, 0xFE = route ends here (-1), 0xFD = lane does not lead on (-2)); refreshed whenever the
// vehicle enters a lane, so that the plan phase and the foe checks read it with one LDS.
__device__ __forceinline__ uint32_t encode_nextlink(const DevScenario& sc, int lane, int k) {
  return k == -1 ? 0xFEu : (k < 0 ? 0xFDu : (uint32_t)(k - __ldg(&sc.lane_rec[lane].link_off)));
}
__device__ __forceinline__ int v_nextlink(const DevScenario& sc, const Tile& t, int i, int lane) {
  uint32_t r = t.meta[i] >> 24;
  return r == 0xFEu ? -1 : (r == 0xFDu ? -2 : __ldg(&sc.lane_rec[lane].link_off) + (int)r);
}
__device__ __forceinline__ int lane_count(const Tile& t, int l) { return (int)t.lane_start[l + 1] - (int)t.lane_start[l]; }

__device__ __forceinline__ void rng4(const Tile& t, uint32_t stream, uint32_t a, uint32_t b, uint32_t out[4]) {
  out[0] = t.env_lo; out[1] = t.env_hi; out[2] = a; out[3] = b;
  philox(out, t.seed_lo ^ (stream * 0x632BE5ABu), t.seed_hi);
}

__device__ __forceinline__ float speed_factor(const Tile& t, int32_t vid, float dev) {
  if (!(dev > 0.0f)) return 1.0f;
  uint32_t r[4];
  rng4(t, STREAM_SF, (uint32_t)vid, 0u, r);
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) sum += (r[i] & 0xFFFFu) + (r[i] >> 16);
  float z = ((float)sum - 262140.0f) * (1.0f / 65536.0f) * 1.2247449f;
  float sf = 1.0f + dev * z;
  return fminf(fmaxf(sf, 0.2f), 2.0f);
}

// which link does a vehicle with (route, cursor) take at the end of `lane`?  -1: route ends, -2: wrong lane.
// The rule (prefer a target lane that is "best", then "ok", then any; first maximum) is tabulated per route step
// and lane index by rs_create (host_choose_link in sim.cu), so this is two dependent loads instead of a loop.
__device__ __forceinline__ int choose_link(const DevScenario& sc, int lane, int route, int cursor) {
  const int k0 = __ldg(&sc.lane_rec[lane].link_off);
  if (__ldg(&sc.lane_rec[lane].internal)) return k0 < __ldg(&sc.lane_rec[lane + 1].link_off) ? k0 : -2;
  const uint32_t r = __ldg(sc.route_step_link + (size_t)(__ldg(sc.route_off + route) + cursor) * 8 + __ldg(&sc.lane_rec[lane].index));
  return r == 0xFEu ? -1 : (r == 0xFDu ? -2 : k0 + (int)r);
}
__device__ __forceinline__ int next_link(const DevScenario& sc, int lane, int route, int cursor) {
  return choose_link(sc, lane, route, cursor);
}

__device__ __forceinline__ int state_now(const DevScenario& sc, const Tile& t, int k) {
  int tl = __ldg(&sc.link_rec[k].tls);
  if (tl < 0) return __ldg(&sc.link_rec[k].state);
  // tls_state[tl] = offset of the current phase's state string (refreshed in shared memory whenever a phase changes)
  return (int)__ldg(sc.state_chars + t.tls_state[tl] + __ldg(&sc.link_rec[k].tlidx));
}

__device__ __forceinline__ bool time_conflict(float seen, float v, float cross, float dist_f, float v_f, float cross_f) {
  float vm = fmaxf(v, 4.0f), vf = fmaxf(v_f, 2.0f);
  float tm_a = seen / vm, tm_l = (seen + cross) / vm;
  float tf_a = dist_f / vf, tf_l = (dist_f + cross_f) / vf;
  return (tf_a < tm_l + 1.0f) && (tf_l + 1.0f > tm_a);
}

// right-of-way: must the vehicle on entry link k wait for one of its foes?
RS_HEAVY bool link_blocked(const DevScenario& sc, const Tile& t, int k, float seen, float v, float cross) {
  const int f0 = __ldg(&sc.link_rec[k].foe_off), f1 = __ldg(&sc.link_rec[k].foe_end);
  if (f0 >= f1) return false;
  const int4* rec = reinterpret_cast<const int4*>(sc.foe_rec);
  int4 a = __ldg(rec + 2 * f0);                          // {link, flags, last_int, from}
  const int4 om = __ldg(reinterpret_cast<const int4*>(&sc.link_rec[k]) + 3);   // {nxt, occ_word, occ_lo, occ_hi}
  const bool windowed = om.y >= 0;
  if (windowed && (((t.occ[om.y] & (uint32_t)om.z) | (t.occ[om.y + 1] & (uint32_t)om.w)) != 0u)) return true;
  for (int fi = f0; fi < f1; ++fi) {
    const int4 c = a;
    a = __ldg(rec + 2 * (fi + 1));                       // next foe's record is in flight while this one is tested
    const int f = c.x, fl = c.y, li = c.z, a0 = c.w;
    if (!windowed && li >= 0 && lane_count(t, li) > 0) return true;   // somebody is crossing my path
    if (!(fl & 1)) continue;                             // I have right of way over f
    const bool occ0 = lane_count(t, a0) > 0;
    if (!occ0 && !(fl & 8)) continue;                    // bit 3: f has a waiting slot inside the junction
    const int4 d = __ldg(rec + 2 * fi + 1);              // {slot, via_len, len_from, len_slot}
    const float via_len = __int_as_float(d.y);
    if (occ0) {
      int h = t.lane_start[a0];
      if (v_nextlink(sc, t, h, a0) == f) {
        float dist_f = __int_as_float(d.z) - t.pos[h];
        int fst = state_now(sc, t, f);
        bool goes = true;
        int hv = v_vtype(t, h);
        if (fst == 'r' || fst == 'u' || fst == 's') goes = false;
        else if (fst == 'y' && dist_f >= brake_gap(t.speed[h], VTT(t, hv, VT_DECEL), 0.0f)) goes = false;
        if (t.speed[h] < kHaltSpeed) goes = false;       // a standing foe is not approaching
        if (goes && time_conflict(seen, v, cross, dist_f, t.speed[h], via_len + VTT(t, hv, VT_LEN))) return true;
      }
    }
    const int a1 = d.x;
    if (a1 >= 0 && lane_count(t, a1) > 0) {
      int h = t.lane_start[a1];
      float dist_f = __int_as_float(d.w) - t.pos[h];
      if (t.speed[h] >= kHaltSpeed &&
          time_conflict(seen, v, cross, dist_f, t.speed[h], via_len - __int_as_float(d.w) + VTT(t, v_vtype(t, h), VT_LEN)))
        return true;
    }
  }
  return false;
}

__device__ __forceinline__ float lane_occ(const Tile& t, int l) {
  float occ = 0.0f;
  for (int i = t.lane_start[l]; i < (int)t.lane_start[l + 1]; ++i) {
    int vt = v_vtype(t, i);
    occ += VTT(t, vt, VT_LEN) + VTT(t, vt, VT_GAP);
  }
  return occ;
}

// stop-line decision for link k, `seen` metres ahead of vehicle i (hop 0 = the link at the end of its lane).
// `binds`: stopping in front of the link would bind the speed now -- right of way and keep-clear of a link further
// ahead are only evaluated then (short lanes are crossed within one tick, so hop 0 alone is not enough).
RS_HEAVY bool must_stop(const DevScenario& sc, const Tile& t, int i, int k, float seen, int hop, int cursor, bool binds) {
  int vt = v_vtype(t, i);
  float len = VTT(t, vt, VT_LEN), decel = VTT(t, vt, VT_DECEL);
  float v = t.speed[i];
  const int4 jn = __ldg(reinterpret_cast<const int4*>(&sc.link_rec[k]) + 5);   // {nxt_link, yield_parent, yield_cross, foe_end}
  const bool from_internal = __ldg(&sc.link_rec[k].from_internal) != 0;
  int yield_link = -1;          // entry link whose foes must be checked (one call site for link_blocked)
  float cross = 0.0f;
  int st = 0;
  if (from_internal) {
    if (!((hop == 0 || binds) && jn.y >= 0)) return false;
    yield_link = jn.y;
    cross = __int_as_float(jn.z) + len;
  } else {
    st = state_now(sc, t, k);
    if (st == 'r' || st == 'u') return true;
    if (st == 'y' || st == 'Y') return seen >= brake_gap(v, decel, 0.0f);
    if (st == 's' && !(v_wait(t, i) > 0 && seen <= 2.0f)) return true;
    if (hop != 0 && !binds) return false;
    bool minor = (st == 'g' || st == 'm' || st == '=' || st == 'Z' || st == 'w' || st == 's' || st == 'o');
    if (!__ldg(&sc.link_rec[k].cont) && minor) { yield_link = k; cross = __ldg(&sc.link_rec[k].via_len) + len; }
  }
  if (yield_link >= 0) {
    bool b = link_blocked(sc, t, yield_link, seen, v, cross);
    if (from_internal) return b;
    if (b) return true;
  }
  const int f0 = __ldg(&sc.link_rec[k].foe_off), f1 = jn.w;
  if (__ldg(&sc.link_rec[k].cont)) {
    // waiting slot inside the junction is taken by a STANDING vehicle (a moving one is simply followed)
    int vl = __ldg(&sc.link_rec[k].via);
    if (lane_count(t, vl) > 0 && t.speed[(int)t.lane_start[vl + 1] - 1] < kHaltSpeed) return true;
  } else if (yield_link < 0) {
    const int4 om = __ldg(reinterpret_cast<const int4*>(&sc.link_rec[k]) + 3);   // {nxt, occ_word, occ_lo, occ_hi}
    if (om.y >= 0) {
      if (((t.occ[om.y] & (uint32_t)om.z) | (t.occ[om.y + 1] & (uint32_t)om.w)) != 0u) return true;
    } else {
      for (int fi = f0; fi < f1; ++fi) {
        int li = __ldg(&sc.foe_rec[fi].last_int);
        if (li >= 0 && lane_count(t, li) > 0) return true;
      }
    }
  }
  // keep the junction clear (SUMO keepClear / getSpaceTillLastStanding): only links with foes, and only when a
  // vehicle was seen beyond the stop line.  Space = room behind the last STANDING vehicle of the lanes ahead (moving
  // vehicles only take their own length), minus the vehicles already inside this junction on my path.
  if (f1 <= f0) return false;
  float need = len + VTT(t, vt, VT_GAP), space = 0.0f;
  bool had = false;
  int cc2 = cursor + 1;
  int route = v_route(t, i);
  int cur;
  cur = __ldg(&sc.link_rec[k].nxt);
  for (int h = 0; h < 3 && __ldg(&sc.lane_rec[cur].internal); ++h) {   // my own path through the junction
    if (lane_count(t, cur) > 0) { had = true; space -= lane_occ(t, cur); }
    int k2 = __ldg(&sc.lane_rec[cur].link_off);
    cur = __ldg(&sc.link_rec[k2].nxt);
  }
  for (int h = 0; h < 6; ++h) {
    int a = t.lane_start[cur], j = (int)t.lane_start[cur + 1] - 1;
    bool stopped = false;
    float lengths = 0.0f;
    if (j >= a) had = true;
    for (; j >= a; --j) {                                          // from the tail forward
      if (t.speed[j] < kHaltSpeed) { stopped = true; break; }
      int yvt = v_vtype(t, j);
      lengths += VTT(t, yvt, VT_LEN) + VTT(t, yvt, VT_GAP);
    }
    if (stopped) { space += (t.pos[j] - VTT(t, v_vtype(t, j), VT_LEN)) - lengths; break; }
    const bool cur_internal = __ldg(&sc.lane_rec[cur].internal) != 0;
    if (!cur_internal) space += __ldg(&sc.lane_rec[cur].len) - lengths;   // junction interiors are no place to stand
    if (space >= need) return false;
    int k2 = next_link(sc, cur, route, cc2);
    if (k2 < 0) return false;
    if (!cur_internal) {
      int st2 = state_now(sc, t, k2);
      if (st2 == 'r' || st2 == 'u' || st2 == 'y') break;
    }
    cur = __ldg(&sc.link_rec[k2].nxt);
    if (!__ldg(&sc.lane_rec[cur].internal)) cc2 += 1;
  }
  return had && space < need;
}

__device__ __forceinline__ int strategic_dir(const DevScenario& sc, int route, int cursor, int lane) {
  int mask = __ldg(sc.route_mask + __ldg(sc.route_off + route) + cursor);
  int okm = mask & 0xFF, bestm = (mask >> 8) & 0xFF, myidx = __ldg(&sc.lane_rec[lane].index);
  int want = !((okm >> myidx) & 1) ? okm : (!((bestm >> myidx) & 1) ? bestm : 0);
  if (!want) return 0;
  for (int d = 1; d < 8; ++d) {
    if (myidx + d < 8 && ((want >> (myidx + d)) & 1)) return 1;
    if (myidx - d >= 0 && ((want >> (myidx - d)) & 1)) return -1;
  }
  return 0;
}

// Plan one vehicle from the state at the start of the tick: next speed + lane it will be in
// laterally (own lane unless a lane change / head swap was decided).
RS_HEAVY void plan_vehicle(const DevScenario& sc, const Tile& t, int i, float& vn_out, int& target_out) {
  int lane = v_lane(t, i);
  int rank = i - (int)t.lane_start[lane];
  int vt = v_vtype(t, i);
  float len = VTT(t, vt, VT_LEN), mingap = VTT(t, vt, VT_GAP), accel = VTT(t, vt, VT_ACCEL);
  float decel = VTT(t, vt, VT_DECEL), tau = VTT(t, vt, VT_TAU);
  float sigma = sc.sigma_override >= 0.0f ? sc.sigma_override : VTT(t, vt, VT_SIGMA);
  float vcapv = VTT(t, vt, VT_VMAX);
  float v = t.speed[i], x = t.pos[i], sf = t.sf[i];
  int route = v_route(t, i), cursor = v_cursor(t, i);
  float lane_len = __ldg(&sc.lane_rec[lane].len);
  float vmaxl = fminf(__ldg(&sc.lane_rec[lane].vmax) * sf, vcapv);
  float vacc = fminf(v + accel, vmaxl);
  float vsafe = vacc, vlead_limit = vacc;
  bool wrong_lane_head = false;
  if (rank > 0) {
    int ld = i - 1, lvt = v_vtype(t, ld);
    float gap = t.pos[ld] - VTT(t, lvt, VT_LEN) - x - mingap;
    vlead_limit = follow_speed(gap, t.speed[ld], VTT(t, lvt, VT_DECEL), decel, tau);
    vsafe = fminf(vsafe, vlead_limit);
  } else {
    float seen = lane_len - x;
    int cur = lane, cc = cursor;
    float la = brake_gap(vacc, decel, 0.0f) + 2.0f * vacc + 5.0f;
    // a lane end further away than the look-ahead distance plus the longest vehicle that could still stick
    // out of the junction cannot bind the speed: no junction logic at all
    const bool far = seen > la + 20.0f;
    int knext = -3;   // link of the NEXT hop when the lane ahead is internal (joined into LinkRec), -3: route lookup
    for (int hop = 0; hop < kMaxHops && !far; ++hop) {
      int k = hop == 0 ? v_nextlink(sc, t, i, lane) : (knext != -3 ? knext : next_link(sc, cur, route, cc));
      if (k == -1) break;
      if (k == -2) {
        vsafe = fminf(vsafe, max_safe_stop_speed(seen, decel, tau));
        if (hop == 0) wrong_lane_head = true;
        break;
      }
      const float vstop = max_safe_stop_speed(seen, decel, tau);
      if (must_stop(sc, t, i, k, seen, hop, cc, vstop < vsafe)) { vsafe = fminf(vsafe, vstop); break; }
      int nxt = __ldg(&sc.link_rec[k].nxt);
      const int4 nx = __ldg(reinterpret_cast<const int4*>(&sc.link_rec[k]) + 4);   // {nxt_len, nxt_vmax, nxt_internal, from_internal}
      vsafe = fminf(vsafe, free_speed(decel, seen, fminf(__int_as_float(nx.y) * sf, vcapv)));
      if (lane_count(t, nxt) > 0) {
        int tl = (int)t.lane_start[nxt + 1] - 1, tvt = v_vtype(t, tl);
        float gap = seen + (t.pos[tl] - VTT(t, tvt, VT_LEN)) - mingap;
        float f = follow_speed(gap, t.speed[tl], VTT(t, tvt, VT_DECEL), decel, tau);
        vsafe = fminf(vsafe, f);
        if (hop == 0) vlead_limit = fminf(vlead_limit, f);
        break;
      }
      seen += __int_as_float(nx.x);
      if (!nx.z) cc += 1;
      cur = nxt;
      knext = __ldg(&sc.link_rec[k].nxt_link);
      if (seen > la) break;
    }
  }
  int left = __ldg(&sc.lane_rec[lane].left), right = __ldg(&sc.lane_rec[lane].right);
  bool internal = __ldg(&sc.lane_rec[lane].internal) != 0;
  // cooperation (LC2013 informFollower analogue): a vehicle of the neighbouring lane that MUST get into this lane
  // (its lane does not lead on) and is urgent becomes a virtual leader for everybody behind it
  if (sc.lane_change && !internal) {
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      const int nl = side == 0 ? left : right;
      if (nl < 0) continue;
      const int a = t.lane_start[nl];
      int j = (int)t.lane_start[nl + 1] - 1;
      float gapu = 0.0f;
      for (; j >= a; --j) {               // from the tail forward: first vehicle that is entirely ahead of me
        gapu = t.pos[j] - VTT(t, v_vtype(t, j), VT_LEN) - x - mingap;
        if (gapu >= 0.0f) break;
      }
      if (j < a) continue;
      const int ur = v_route(t, j), uc = v_cursor(t, j), uvt = v_vtype(t, j);
      const int masku = __ldg(sc.route_mask + __ldg(sc.route_off + ur) + uc);
      if ((masku >> __ldg(&sc.lane_rec[nl].index)) & 1) continue;                      // its lane leads on: not urgent
      if (!(__ldg(&sc.lane_rec[nl].len) - t.pos[j] < 60.0f || v_wait(t, j) > 3)) continue;
      const int du = strategic_dir(sc, ur, uc, nl);
      if ((du > 0 ? __ldg(&sc.lane_rec[nl].left) : (du < 0 ? __ldg(&sc.lane_rec[nl].right) : -1)) != lane) continue;
      if (!(__ldg(&sc.lane_rec[lane].perm) & __ldg(sc.vtype_bit + uvt))) continue;
      vsafe = fminf(vsafe, follow_speed(gapu - kCoopMargin, t.speed[j], VTT(t, uvt, VT_DECEL), decel, tau));
    }
  }
  float vmin_n = fmaxf(0.0f, v - decel);
  float vmin_e = fmaxf(0.0f, v - fmaxf(decel, kEmergencyDecel));
  float vmin = fminf(vmin_n, fmaxf(vsafe, vmin_e));
  float vcand = fmaxf(vmin, vsafe);
  float vn = vcand;
  if (sigma > 0.0f) {
    uint32_t r[4];
    rng4(t, STREAM_DAWDLE, (uint32_t)t.vid[i], (uint32_t)t.tick, r);
    float xi = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
    vn = fmaxf(vmin, dawdle(vcand, accel, sigma, xi));
  }
  int target = -1;
  if (sc.lane_change && !internal && (left >= 0 || right >= 0) && x + vn <= lane_len) {
    int mask = __ldg(sc.route_mask + __ldg(sc.route_off + route) + cursor);
    int okm = mask & 0xFF, bestm = (mask >> 8) & 0xFF, myidx = __ldg(&sc.lane_rec[lane].index);
    int dir = strategic_dir(sc, route, cursor, lane);
    // strategic (must): the route cannot continue from this lane.  A lane that leads on but is not "best" only
    // makes the best lanes attractive (no speed loss needed to go there); any lane that leads on may be used to get
    // around a blocked leader.
    const bool strategic = dir != 0 && !((okm >> myidx) & 1);
    const bool cur_best = (bestm >> myidx) & 1;
    // urgent: the route cannot continue from this lane and the lane end is near (or the vehicle already
    // stands): accept any gap the neighbours can still handle with emergency braking
    const bool urgent = strategic && (lane_len - x < 60.0f || v_wait(t, i) > 3);
    int vbit = __ldg(sc.vtype_bit + vt);
    for (int pass = 0; pass < 2; ++pass) {
      int d;
      if (strategic) { if (pass) break; d = dir; }
      else {
        if (v_lcc(t, i) > 0 || (cur_best && !(vlead_limit < vacc - 1.0f))) break;
        d = pass == 0 ? 1 : -1;
      }
      if (d == 0) break;
      if ((d > 0) == ((t.tick & 1) != 0)) continue;   // even ticks: leftward, odd ticks: rightward
      int nl = d > 0 ? left : right;
      if (nl < 0 || !(__ldg(&sc.lane_rec[nl].perm) & vbit)) continue;
      const int nlidx = __ldg(&sc.lane_rec[nl].index);
      if (!strategic && !((okm >> nlidx) & 1)) continue;
      int a = t.lane_start[nl], b = t.lane_start[nl + 1], j = a;
      while (j < b && t.pos[j] >= x) ++j;
      float vfol = vacc;
      bool ok = true;
      if (j > a) {
        int ld = j - 1, lvt = v_vtype(t, ld);
        float gap = t.pos[ld] - VTT(t, lvt, VT_LEN) - x - mingap;
        if (gap < 0.0f) ok = false;
        else {
          vfol = follow_speed(gap, t.speed[ld], VTT(t, lvt, VT_DECEL), decel, tau);
          if (vfol < v - (urgent ? fmaxf(decel, kEmergencyDecel) : decel)) ok = false;
        }
      }
      if (ok && j < b) {
        int fvt = v_vtype(t, j);
        float gap = x - len - t.pos[j] - VTT(t, fvt, VT_GAP);
        if (gap < 0.0f) ok = false;
        else if (urgent) {
          if (gap < brake_gap(t.speed[j], fmaxf(VTT(t, fvt, VT_DECEL), kEmergencyDecel), 1.0f)) ok = false;   // 1 s: it reacts a tick late
        } else {
          float vf = follow_speed(gap, v, decel, VTT(t, fvt, VT_DECEL), VTT(t, fvt, VT_TAU));
          if (vf < t.speed[j] + VTT(t, fvt, VT_ACCEL) - VTT(t, fvt, VT_DECEL)) ok = false;
        }
      }
      // nobody behind in the target lane: the follower may still be upstream of it, about to come out of a junction
      if (ok && j >= b && x - len < 60.0f) {
        const int w1 = __ldg(sc.lane_watch_off + nl + 1);
        for (int w = __ldg(sc.lane_watch_off + nl); ok && w < w1; ++w) {
          const int pl = __ldg(sc.lane_watch_lane + w);
          if (lane_count(t, pl) == 0) continue;
          const int h = t.lane_start[pl], hvt = v_vtype(t, h);
          const float gap = (x - len) + __ldg(sc.lane_watch_dist + w) + (__ldg(&sc.lane_rec[pl].len) - t.pos[h]) - VTT(t, hvt, VT_GAP);
          bool unsafe;
          if (urgent) unsafe = gap < brake_gap(t.speed[h], fmaxf(VTT(t, hvt, VT_DECEL), kEmergencyDecel), 1.0f);
          else {
            const float vf = follow_speed(gap, v, decel, VTT(t, hvt, VT_DECEL), VTT(t, hvt, VT_TAU));
            unsafe = vf < t.speed[h] + VTT(t, hvt, VT_ACCEL) - VTT(t, hvt, VT_DECEL);
          }
          if (!unsafe) continue;
          int cur = pl, cc = v_cursor(t, h);      // does its route lead onto the target lane?
          const int hr = v_route(t, h);
          for (int hop = 0; hop < 4; ++hop) {
            const int k = hop == 0 ? v_nextlink(sc, t, h, pl) : next_link(sc, cur, hr, cc);
            if (k < 0) break;
            const int nxt = __ldg(&sc.link_rec[k].nxt);
            if (nxt == nl) { ok = false; break; }
            if (!__ldg(&sc.lane_rec[nxt].internal)) cc += 1;
            cur = nxt;
          }
        }
      }
      if (!ok) continue;
      if (!strategic) {   // required speed gain: none towards a best lane, 1 m/s between best lanes, 2 m/s away from them
        const bool nl_best = (bestm >> nlidx) & 1;
        const float gain = fminf(vfol, vacc) - vlead_limit;
        if (nl_best && !cur_best) { if (!(gain >= -0.5f)) continue; }
        else if (nl_best) { if (!(gain > 1.0f)) continue; }
        else if (!(gain > 2.0f && vlead_limit < vacc - 1.0f)) continue;
      }
      target = nl;
      vn = fmaxf(0.0f, fminf(vn, vfol));
      break;
    }
  }
  // deadlock breaker: two standing lane heads that each need the other's lane trade places
  if (target < 0 && wrong_lane_head && sc.lane_change && v < kHaltSpeed && lane_len - x < 1.0f) {
    int d = strategic_dir(sc, route, cursor, lane);
    int nl = d > 0 ? left : (d < 0 ? right : -1);
    if (nl >= 0 && lane_count(t, nl) > 0 && (__ldg(&sc.lane_rec[nl].perm) & __ldg(sc.vtype_bit + vt))) {
      int y = t.lane_start[nl];
      if (t.speed[y] < kHaltSpeed && __ldg(&sc.lane_rec[nl].len) - t.pos[y] < 1.0f &&
          (__ldg(&sc.lane_rec[lane].perm) & __ldg(sc.vtype_bit + v_vtype(t, y))) &&
          v_nextlink(sc, t, y, nl) == -2 &&
          strategic_dir(sc, v_route(t, y), v_cursor(t, y), nl) == -d) {
        target = nl;
        vn = 0.0f;
      }
    }
  }
  vn_out = vn;
  target_out = target >= 0 ? target : lane;
}

// ---------------------------------------------------------------------------------------------
// TMA 1-D bulk copies (cp.async.bulk, SASS: UBLKCP) + mbarrier transaction tracking
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// block-wide exclusive prefix sum over `n` ints in shared memory (in: cnt, out: start u16[n+1])
template <int BLOCK>
__device__ void block_prefix(const int32_t* cnt, uint16_t* start, int n, int32_t* warp_tot) {
  const int tid = threadIdx.x % BLOCK, lane = tid & 31, wid = tid >> 5;   // BLOCK = threads per instance
  const int per = (n + BLOCK - 1) / BLOCK;
  const int a = min(tid * per, n), b = min(a + per, n);
  int s = 0;
  for (int i = a; i < b; ++i) s += cnt[i] & 0x3FFFFFFF;
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < wid; ++w) base += warp_tot[w];
  int run = base + inc - s;
  for (int i = a; i < b; ++i) { start[i] = (uint16_t)run; run += cnt[i] & 0x3FFFFFFF; }
  if (tid == BLOCK - 1) start[n] = (uint16_t)run;
  __syncthreads();
}

}  // namespace rs
