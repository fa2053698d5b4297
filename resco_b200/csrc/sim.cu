// sim.cu -- env-step kernel, auxiliary kernels and the C-ABI (include/resco_b200.h).
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
// There is NO CPU fallback: every entry point fails with RS_ERR_NODEVICE without a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sim_kernels.cuh"
#include "agents.cuh"

namespace rs {

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up (host + device agree through this struct)
struct SmemLayout {
  int vcap, L, n_tls, S, O, SL, n_vt;
  int single;   // one tile buffer: the per-tick re-sort goes through registers (vcap <= 2 x threads per instance)
  int gmem;     // the per-vehicle region lives in global memory (DevSim::workspace); offsets below it restart at 0
  // per-vehicle region (offsets from the vehicle-region base: shared memory, or the CTA's workspace slot)
  size_t off_bufA, off_bufB, off_vn, off_newlane, off_newidx, off_mnext, off_arr, off_dirty;
  size_t veh_total;
  // per-lane / per-signal / per-origin region (offsets from the instance's shared-memory base)
  size_t off_lane_start, off_start2, off_cnt2, off_mhead;
  size_t off_tls_phase, off_tls_end, off_tls_state, off_next_phase, off_origin_cur, off_origin_backlog, off_origin_cand;
  size_t off_vt, off_hdr, off_misc, off_obs, off_mbar, off_oklist, off_occ;
  size_t total;   // shared memory per instance
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline SmemLayout make_layout_ex(const DevScenario& sc, int tile_cap, int single, int gmem) {
  SmemLayout m;
  m.single = single; m.gmem = gmem;
  m.vcap = tile_cap; m.L = sc.n_lanes; m.n_tls = sc.n_tls; m.S = sc.n_signals; m.O = sc.n_origins;
  m.SL = sc.n_sig_lanes; m.n_vt = sc.n_vtypes;
  size_t o = 0;
  m.off_bufA = o; o = align16(o + (size_t)kVehWords * m.vcap * 4);
  m.off_bufB = m.single ? m.off_bufA : o; if (!m.single) o = align16(o + (size_t)kVehWords * m.vcap * 4);
  m.off_vn = o; o = align16(o + (size_t)m.vcap * 4);
  m.off_newlane = o; o = align16(o + (size_t)m.vcap * 2);
  m.off_newidx = o; o = align16(o + (size_t)m.vcap * 2);
  m.off_mnext = o; o = align16(o + (size_t)m.vcap * 2);
  m.off_arr = o; o = align16(o + (size_t)m.vcap * 2);
  m.off_dirty = o; o = align16(o + ((size_t)2 * m.vcap + (size_t)(m.O > 0 ? m.O : 1)) * 2);   // lanes touched this tick
  m.veh_total = o;
  if (m.gmem) o = 0;
  m.off_lane_start = o; o = align16(o + (size_t)(m.L + 1) * 2);
  m.off_start2 = o; o = align16(o + (size_t)(m.L + 1) * 2);
  m.off_cnt2 = o; o = align16(o + (size_t)m.L * 4);
  m.off_mhead = o; o = align16(o + (size_t)m.L * 4);
  m.off_tls_phase = o; o = align16(o + (size_t)m.n_tls * 4);
  m.off_tls_end = o; o = align16(o + (size_t)m.n_tls * 4);
  m.off_tls_state = o; o = align16(o + (size_t)m.n_tls * 4);
  m.off_next_phase = o; o = align16(o + (size_t)(m.S > 0 ? m.S : 1) * 4);
  m.off_origin_cur = o; o = align16(o + (size_t)(m.O > 0 ? m.O : 1) * 4);
  m.off_origin_backlog = o; o = align16(o + (size_t)(m.O > 0 ? m.O : 1) * 4);
  m.off_origin_cand = o; o = align16(o + (size_t)(m.O > 0 ? m.O : 1) * 16);
  m.off_vt = o; o = align16(o + (size_t)m.n_vt * 8 * 4);
  m.off_hdr = o; o = align16(o + (size_t)kHdrInts * 4);
  m.off_misc = o; o = align16(o + 48 * 4);
  // the per-lane observation scratch is only live in observe_body: it shares the plan scratch (vn + newlane, which
  // are contiguous) when that is in shared memory and it fits there; ~0 = "lives at off_vn of the vehicle region"
  if (!m.gmem && (size_t)m.SL * 5 * 4 <= (size_t)m.vcap * 6) m.off_obs = ~(size_t)0;
  else { m.off_obs = o; o = align16(o + (size_t)(m.SL > 0 ? m.SL : 1) * 5 * 4); }
  m.off_mbar = o; o = align16(o + 16);
  m.off_oklist = o; o = align16(o + (size_t)(m.O > 0 ? m.O : 1) * 2);                          // origins with a candidate
  m.off_occ = o; o = align16(o + ((size_t)(m.L + 31) / 32 + 2) * 4);                          // lane-occupancy bits
  m.total = o;
  return m;
}
__host__ __device__ inline SmemLayout make_layout(const DevScenario& sc) {
  return make_layout_ex(sc, sc.tile_cap, sc.tile_single, sc.tile_gmem);
}

// Optional per-phase cycle accounting (build with -DRS_PHASE_CLOCKS=1, tools/phase_clocks.sh): thread 0 of the CTA
// charges the cycles since the previous mark to slot i; k_run adds the CTA totals to DevSim::phase_clocks.
#ifndef RS_PHASE_CLOCKS
#define RS_PHASE_CLOCKS 0
#endif
#if RS_PHASE_CLOCKS
__shared__ long long s_pclk[24];
__shared__ long long s_pclk_last;
#define PCLK(i) do { if (threadIdx.x == 0) { long long c_ = clock64(); s_pclk[i] += c_ - s_pclk_last; s_pclk_last = c_; } } while (0)
#else
#define PCLK(i) do { } while (0)
#endif
enum { PC_STAGE = 0, PC_S0, PC_S1, PC_S2, PC_S3A, PC_S3B, PC_S4, PC_S5, PC_S6, PC_S7, PC_OBS, PC_WRITE, PC_SCHED, PC_N };

// misc slots
enum { M_NARR = 0, M_NOK, M_NAFTER, M_NDIRTY, M_NOKC, M_MAYDEFER /* outgrowing the tile defers the instance instead of refusing insertions */, M_BAIL /* the instance outgrew the tile: its step is redone by the overflow pass */, M_WARP = 16 /* 32 ints of warp totals */ };

#ifndef RS_PLAN_SPREAD
#define RS_PLAN_SPREAD 1
#endif

constexpr int kDirty = 0x40000000;   // flag bit in cnt2[l]: the lane gained or lost a vehicle this tick

// ok_dd: -1 not ok, -2 refused by capacity, else depart delay; rank: vehicles of the lane ahead of the newcomer
// (departPos="random_free"), kRankBack = it goes to the back of the lane (departPos="base"); pos: its front position
struct OriginCand { int32_t vid; uint16_t route; int16_t ok_dd; uint16_t vt; uint16_t rank; float pos; };
static_assert(sizeof(OriginCand) == 16, "OriginCand is 16 bytes in the shared-memory layout");
constexpr uint32_t kRankBack = 0xFFFFu;

// ------------------------------------------------------------------------------------------------
template <int BLOCK, int G>
__device__ __forceinline__ void tick_body(const DevSim& D, const SmemLayout& m, unsigned char* smem, unsigned char* vb, Tile& T,
                          uint32_t*& cur, uint32_t*& oth, uint16_t*& start2, const int env) {
  const DevScenario& sc = D.sc;
  const int tid = threadIdx.x % BLOCK;   // BLOCK = threads per instance (a CTA may hold several instances)
  const int L = m.L;
  float* vn = (float*)(vb + m.off_vn);               // vb: base of the per-vehicle region (shared memory or workspace)
  uint16_t* newlane = (uint16_t*)(vb + m.off_newlane);
  uint16_t* newidx = (uint16_t*)(vb + m.off_newidx);
  uint16_t* mnext = (uint16_t*)(vb + m.off_mnext);
  uint16_t* arr = (uint16_t*)(vb + m.off_arr);
  int32_t* cnt2 = (int32_t*)(smem + m.off_cnt2);     // per-lane vehicle count, kept current across ticks
  uint16_t* dirty = (uint16_t*)(vb + m.off_dirty);
  uint16_t* oklist = (uint16_t*)(smem + m.off_oklist);
  int32_t* mhead = (int32_t*)(smem + m.off_mhead);
  int32_t* hdr = (int32_t*)(smem + m.off_hdr);
  int32_t* misc = (int32_t*)(smem + m.off_misc);
  int32_t* origin_cur = (int32_t*)(smem + m.off_origin_cur);
  int32_t* origin_backlog = (int32_t*)(smem + m.off_origin_backlog);
  OriginCand* cand = (OriginCand*)(smem + m.off_origin_cand);
  const int n = hdr[H_NVEH];

  // ---- S0: traffic lights (static-program countdown), per-lane counters ----
  for (int t = tid; t < m.n_tls; t += BLOCK) {
    int p0 = __ldg(sc.tls_phase_off + t), np = __ldg(sc.tls_phase_off + t + 1) - p0;
    int guard = 0;
    while (T.tick >= T.tls_end[t] && guard++ < 64) {
      int ph = (T.tls_phase[t] + 1) % np;
      T.tls_phase[t] = ph;
      int d = __ldg(sc.phase_dur + p0 + ph);
      T.tls_end[t] += d > 0 ? d : 1;
    }
    T.tls_state[t] = __ldg(sc.phase_state_off + p0 + T.tls_phase[t]);
  }
  if (tid == 0) { misc[M_NARR] = 0; misc[M_NOK] = 0; misc[M_NDIRTY] = 0; misc[M_NOKC] = 0; }
  __syncthreads();
  PCLK(PC_S0);

  // a lane that gains or loses a vehicle is put on the dirty list once (flag bit in its counter)
  auto mark_dirty = [&](int l) {
    int old = atomicOr(&cnt2[l], kDirty);
    if (!(old & kDirty)) { int s = atomicAdd(&misc[M_NDIRTY], 1); dirty[s] = (uint16_t)l; }
  };

  // ---- S1: plan (reads only start-of-tick state) ----
  // Two-warp instances: the last, partial pass over the vehicles is split evenly over the two warps (each takes a run of
  // consecutive vehicles) instead of filling the first and leaving the second idle: a warp's time in this phase is the
  // SUM of its lanes' junction look-aheads (they diverge from each other), so the phase ends when the fuller warp does
  // (cologne8, ~78 vehicles per instance: 39 + 39 instead of 46 + 32; +2 % in the driver window, +3 % over whole episodes).
  // 16-warp instances do not gain (ingolstadt21: -1 %; dealt with a stride of 16 instead of in runs: -13 %).
  {
    [[maybe_unused]] constexpr int W = BLOCK / 32;
    for (int base = 0; base < n; base += BLOCK) {
      const int r = n - base;
      int i = base + tid;
#if RS_PLAN_SPREAD
      if (W == 2 && r < BLOCK) {
        const int chunk = (r + W - 1) / W, k = (tid >> 5) * chunk + (tid & 31);
        i = ((tid & 31) < chunk && k < r) ? base + k : n;
      }
#endif
      if (i < n) {
        float v; int tg;
        plan_vehicle(sc, T, i, v, tg);
        vn[i] = v; newlane[i] = (uint16_t)tg;
      }
    }
  }
  __syncthreads();
  PCLK(PC_S1);

  // ---- S2: move: update in place, hand-off across lanes, bucket movers by target lane ----
  for (int i = tid; i < n; i += BLOCK) {
    int l = v_lane(T, i);
    int vt = v_vtype(T, i);
    float v1 = vn[i];
    float vmaxl = fminf(__ldg(&sc.lane_rec[l].vmax) * T.sf[i], VTT(T, vt, VT_VMAX));
    T.speed[i] = v1;
    uint32_t w = T.wr[i];
    uint32_t wait = v1 < kHaltSpeed ? (w & 0xFFFFu) + 1u : 0u;
    T.wr[i] = (w & 0xFFFF0000u) | (wait & 0xFFFFu);
    if (v1 < kHaltSpeed) { const uint32_t aw = T.aw[i]; if ((aw & 0xFFFFu) != 0xFFFFu) T.aw[i] = aw + 1u; }
    T.tloss[i] += (vmaxl - v1) / vmaxl;
    int lcc = v_lcc(T, i);
    if (lcc > 0) lcc -= 1;
    float p = T.pos[i] + v1;
    int curl = l, cc = v_cursor(T, i);
    int route = v_route(T, i);
    int tg = newlane[i];
    if (tg != l) { curl = tg; lcc = kLcCooldown; }
    else {
      int guard = 0;
      while (p > __ldg(&sc.lane_rec[curl].len) && guard++ < 64) {
        int k = guard == 1 ? v_nextlink(sc, T, i, l) : next_link(sc, curl, route, cc);
        if (k == -1) { curl = -1; break; }
        if (k == -2) { p = __ldg(&sc.lane_rec[curl].len); break; }
        p -= __ldg(&sc.lane_rec[curl].len);
        curl = __ldg(&sc.link_rec[k].nxt);
        if (!__ldg(&sc.lane_rec[curl].internal)) cc += 1;
      }
    }
    T.pos[i] = p;
    T.rc[i] = (T.rc[i] & 0xFFFFu) | ((uint32_t)cc << 16);
    uint32_t mt = (T.meta[i] & 0xFFFF00FFu) | ((uint32_t)lcc << 8);
    if (curl >= 0 && curl != l) mt = (mt & 0x00FFFFFFu) | (encode_nextlink(sc, curl, next_link(sc, curl, route, cc)) << 24);
    T.meta[i] = mt;
    if (curl < 0) {
      newlane[i] = (uint16_t)kArrived;
      if (D.trip_rec && !misc[M_BAIL])   // tripinfo record (multi_signal.py:127-129); a deferred instance's step is redone
        D.trip_rec[(size_t)env * sc.n_trips + T.vid[i]] =
            make_int4(T.tick, (int)(T.ed[i] >> 16), __float_as_int(T.tloss[i]), (int)(T.dl[i] & 0xFFFFu));
      if (D.trip_rec && !misc[M_BAIL]) D.trip_wait[(size_t)env * sc.n_trips + T.vid[i]] = (float)(T.aw[i] & 0xFFFFu);
      atomicAdd(&cnt2[l], -1);
      mark_dirty(l);
      int s = atomicAdd(&misc[M_NARR], 1);
      arr[s] = (uint16_t)i;
    } else {
      newlane[i] = (uint16_t)curl;
      if (curl != l) {
        atomicAdd(&cnt2[l], -1);
        atomicAdd(&cnt2[curl], 1);
        mark_dirty(l); mark_dirty(curl);
        int old = atomicExch(&mhead[curl], i);
        mnext[i] = (uint16_t)(old < 0 ? 0xFFFF : old);
      }
    }
  }
  __syncthreads();
  PCLK(PC_S2);

  // ---- S3a: arrivals (deterministic CSR order) and insertion candidates ----
  if (tid == 0) {
    int na = misc[M_NARR];
    misc[M_NAFTER] = n - na;
    float sd = __int_as_float(hdr[H_F_DELAY_ARR]), sdur = __int_as_float(hdr[H_F_DUR_ARR]), swait = __int_as_float(hdr[H_F_WAIT_ARR]);
    int last = -1;
    for (int k = 0; k < na; ++k) {
      int best = 0x7FFFFFFF;
      for (int q = 0; q < na; ++q) { int a = arr[q]; if (a > last && a < best) best = a; }
      last = best;
      sd += T.tloss[best] + (float)(T.dl[best] & 0xFFFFu);
      sdur += (float)(T.tick - (int)(T.ed[best] >> 16));
      swait += (float)(T.aw[best] & 0xFFFFu);
    }
    hdr[H_F_DELAY_ARR] = __float_as_int(sd); hdr[H_F_DUR_ARR] = __float_as_int(sdur); hdr[H_F_WAIT_ARR] = __float_as_int(swait);
    hdr[H_NARR] += na;
  }
  for (int o = tid; o < m.O; o += BLOCK) {
    // trip-table demand: origin_backlog[o] holds the departure time of the origin's next trip (float bits),
    // so an origin with nothing due costs one shared-memory compare
    if (!sc.synthetic && !(__int_as_float(origin_backlog[o]) <= (float)T.tick)) { cand[o].ok_dd = -1; continue; }
    int lane = __ldg(sc.origin_lane + o);
    OriginCand c; c.ok_dd = -1; c.route = 0; c.vt = 0; c.vid = 0; c.rank = (uint16_t)kRankBack; c.pos = 0.0f;
    bool have = false, rnd = false;
    int dd = 0;
    if (sc.synthetic) {
      uint32_t r[4];
      rng4(T, STREAM_DEMAND, (uint32_t)o, (uint32_t)T.tick, r);
      if ((int32_t)(r[0] >> 8) < __ldg(sc.origin_rate + o)) origin_backlog[o] += 1;
      if (origin_backlog[o] > 0) {
        int r0 = __ldg(sc.origin_route_off + o), nr = __ldg(sc.origin_route_off + o + 1) - r0;
        if (nr <= 0) origin_backlog[o] = 0;
        else {
          rng4(T, STREAM_ROUTE, (uint32_t)o, (uint32_t)origin_cur[o], r);
          c.route = __ldg(sc.origin_route + r0 + (int)(r[0] % (uint32_t)nr));
          c.vt = (uint16_t)sc.synthetic_vtype;
          c.vid = (o << 16) | (origin_cur[o] & 0xFFFF);
          have = true;
        }
      }
    } else {
      int ci = __ldg(sc.origin_off + o) + origin_cur[o];
      if (ci < __ldg(sc.origin_off + o + 1) && !(__ldg(sc.trip_depart + ci) > (float)T.tick)) {
        c.route = __ldg(sc.trip_route + ci); c.vt = (uint16_t)__ldg(sc.trip_vtype + ci); c.vid = ci;
        dd = T.tick - (int)__ldg(sc.trip_depart + ci);
        rnd = __ldg(sc.trip_depart_pos + ci) == 1;
        have = true;
      }
    }
    if (have) {
      const float len = VTT(T, c.vt, VT_LEN), mingap = VTT(T, c.vt, VT_GAP);
      const float lane_len = __ldg(&sc.lane_rec[lane].len);
      const int la = T.lane_start[lane], lb = T.lane_start[lane + 1];
      // upstream safety: nobody who is about to drive onto this lane may be forced into hard braking by a vehicle
      // standing with its back `back` metres into the lane
      auto upstream_clear = [&](float back) {
        for (int w = __ldg(sc.origin_watch_off + o); w < __ldg(sc.origin_watch_off + o + 1); ++w) {
          int pl = __ldg(sc.origin_watch_lane + w);
          // post-move head of pl: first stayer of the old segment vs. the front-most mover into pl
          int a = T.lane_start[pl], b = T.lane_start[pl + 1];
          int hs = -1;
          for (int i = a; i < b; ++i) if (newlane[i] == pl) { hs = i; break; }
          int hm = -1;
          for (int q = mhead[pl]; q >= 0; q = (mnext[q] == 0xFFFF ? -1 : (int)mnext[q]))
            if (hm < 0 || T.pos[q] > T.pos[hm] || (T.pos[q] == T.pos[hm] && q < hm)) hm = q;
          int h = hs;
          if (hm >= 0 && (hs < 0 || T.pos[hm] > T.pos[hs])) h = hm;
          if (h < 0) continue;
          // cheap test first: only a head that could not brake in time is worth the route look-ahead
          int hvt = v_vtype(T, h);
          float gap = ((__ldg(&sc.lane_rec[pl].len) - T.pos[h]) + __ldg(sc.origin_watch_dist + w) - VTT(T, hvt, VT_GAP)) + back;
          if (!(gap < brake_gap(T.speed[h], VTT(T, hvt, VT_DECEL), VTT(T, hvt, VT_TAU)))) continue;
          int cur = pl, cc = v_cursor(T, h), hr = v_route(T, h);
          bool reaches = false;
          for (int hop = 0; hop < 4; ++hop) {
            int k = hop == 0 ? v_nextlink(sc, T, h, pl) : next_link(sc, cur, hr, cc);
            if (k < 0) break;
            int nxt = __ldg(&sc.link_rec[k].nxt);
            if (nxt == lane) { reaches = true; break; }
            if (!__ldg(&sc.lane_rec[nxt].internal)) cc += 1;
            cur = nxt;
          }
          if (reaches) return false;
        }
        return true;
      };
      bool ok = false;
      c.pos = len;
      // departPos="random_free": ten uniformly drawn positions are tried for one where the vehicle fits between the
      // post-move content of the lane (stayers of the old segment + movers chained at mhead), then the base rule
      if (rnd && len <= lane_len) {
        for (int k = 0; k < 10 && !ok; ++k) {
          uint32_t r[4];
          rng4(T, STREAM_DEPARTPOS, (uint32_t)c.vid, (uint32_t)T.tick * 4u + (uint32_t)(k >> 2), r);
          const float u = (float)(r[k & 3] >> 8) * (1.0f / 16777216.0f);
          const float p = len + u * (lane_len - len);
          const float back = p - len;
          bool fits = true; int ahead = 0, behind = 0;
          auto look = [&](int i) {
            const float x = T.pos[i]; const int xvt = v_vtype(T, i);
            if (x >= p) { ahead += 1; if (x - VTT(T, xvt, VT_LEN) - p - mingap < 0.0f) fits = false; }
            else { behind += 1; if (back - x - VTT(T, xvt, VT_GAP) < brake_gap(T.speed[i], VTT(T, xvt, VT_DECEL), VTT(T, xvt, VT_TAU))) fits = false; }
          };
          for (int i = la; i < lb; ++i) if (newlane[i] == lane) look(i);
          for (int q = mhead[lane]; q >= 0; q = (mnext[q] == 0xFFFF ? -1 : (int)mnext[q])) look(q);
          if (fits && behind == 0) fits = upstream_clear(back);
          if (fits) { ok = true; c.rank = (uint16_t)ahead; c.pos = p; }
        }
      }
      if (!ok) {
        ok = len <= lane_len;
        if (ok) {
          // post-move tail of the origin lane = last element of the merged (stayers + movers) order
          int ls = -1;
          for (int i = lb - 1; i >= la; --i) if (newlane[i] == lane) { ls = i; break; }
          int tailv = ls;
          int best = -1;   // mover with the smallest (pos, then largest idx)
          for (int q = mhead[lane]; q >= 0; q = (mnext[q] == 0xFFFF ? -1 : (int)mnext[q]))
            if (best < 0 || T.pos[q] < T.pos[best] || (T.pos[q] == T.pos[best] && q > best)) best = q;
          if (best >= 0 && (ls < 0 || !(T.pos[best] > T.pos[ls]))) tailv = best;
          if (tailv >= 0) {
            int tvt = v_vtype(T, tailv);
            if (T.pos[tailv] - VTT(T, tvt, VT_LEN) - len - mingap < 0.0f) ok = false;
          }
        }
        if (ok) ok = upstream_clear(0.0f);
      }
      if (ok) { c.ok_dd = dd; int s = atomicAdd(&misc[M_NOKC], 1); oklist[s] = (uint16_t)o; atomicAdd(&misc[M_NOK], 1); }
    }
    cand[o] = c;
  }
  __syncthreads();
  PCLK(PC_S3A);

  // ---- S3b: capacity resolution (origin order) ----
  const int nokc = misc[M_NOKC];
  {
    const int n_after = misc[M_NAFTER], n_ok = misc[M_NOK];
    const bool all = n_after + n_ok <= m.vcap;
    // the TILE is outgrown, not the store: defer the instance instead of refusing insertions (a slot that was skipped for the
    // heavy list stays skipped)
    if (!all && misc[M_MAYDEFER] && tid == 0 && misc[M_BAIL] == 0) misc[M_BAIL] = 1;
    for (int j = tid; j < nokc; j += BLOCK) {
      const int o = oklist[j];
      if (cand[o].ok_dd < 0) continue;
      bool acc = all;
      if (!all) {
        int before = 0;
        for (int q = 0; q < o; ++q) before += cand[q].ok_dd >= 0 || cand[q].ok_dd == -2;
        acc = n_after + before < m.vcap;
      }
      if (acc) {
        const int lane = __ldg(sc.origin_lane + o);
        atomicAdd(&cnt2[lane], 1);
        mark_dirty(lane);
        atomicAdd(&hdr[H_NINS], 1);
        origin_cur[o] += 1;
        if (sc.synthetic) origin_backlog[o] -= 1;
        else {
          const int ci = __ldg(sc.origin_off + o) + origin_cur[o];
          origin_backlog[o] = __float_as_int(ci < __ldg(sc.origin_off + o + 1) ? __ldg(sc.trip_depart + ci) : 3.0e38f);
        }
      } else {   // refused by the capacity of the store: counted (RsStats.n_cap_refused), still "ok before" for later origins;
                 // a deferred instance's counters are never written back
        cand[o].ok_dd = -2; atomicAdd(&hdr[H_NREF], 1);
      }
    }
  }
  __syncthreads();
  PCLK(PC_S3B);

  // ---- S4: new lane offsets ----
  block_prefix<BLOCK>(cnt2, start2, L, misc + M_WARP);   // counts are read without the dirty flag bit
  PCLK(PC_S4);

  // ---- S5: per-lane merge: stayers keep their order, movers merge in by position ----
  const int ndirty = misc[M_NDIRTY];
  for (int dj = tid; dj < ndirty; dj += BLOCK) {
    const int l = dirty[dj];
    int a = T.lane_start[l], b = T.lane_start[l + 1];
    int head = mhead[l];
    int w = start2[l];
    // next mover in (pos desc, idx asc) order after (ppos, pidx)
    float ppos = 3.0e38f; int pidx = -1;
    auto next_mover = [&](float pp, int pi) {
      int best = -1;
      for (int q = head; q >= 0; q = (mnext[q] == 0xFFFF ? -1 : (int)mnext[q])) {
        float xq = T.pos[q];
        bool after = (xq < pp) || (xq == pp && q > pi);
        if (!after) continue;
        if (best < 0 || xq > T.pos[best] || (xq == T.pos[best] && q < best)) best = q;
      }
      return best;
    };
    // a newcomer placed between the vehicles of the lane (departPos="random_free") keeps the slot of its rank free
    int hole = -1;
    for (int j = 0; j < nokc; ++j) {
      const int o = oklist[j];
      if (cand[o].ok_dd >= 0 && cand[o].rank != kRankBack && __ldg(sc.origin_lane + o) == l) hole = w + (int)cand[o].rank;
    }
    int mv = head >= 0 ? next_mover(ppos, pidx) : -1;
    for (int i = a; i < b; ++i) {
      if (newlane[i] != l) continue;
      while (mv >= 0 && T.pos[mv] > T.pos[i]) { w += (w == hole); newidx[mv] = (uint16_t)w++; mv = next_mover(T.pos[mv], mv); }
      w += (w == hole);
      newidx[i] = (uint16_t)w++;
    }
    while (mv >= 0) { w += (w == hole); newidx[mv] = (uint16_t)w++; mv = next_mover(T.pos[mv], mv); }
  }
  __syncthreads();
  PCLK(PC_S5);

  // ---- S6: scatter into the other buffer; newcomers at the back of their origin lane ----
  {
    Tile U = T; tile_bind(U, oth, m.vcap);    // single-buffer mode: oth == cur, the permutation goes through registers
    // (compiled for the 512-thread shapes only: the small-tile shapes keep the ping-pong tile and their register budget)
    if (BLOCK >= 512 && m.single) {
      uint32_t w[2][kVehWords]; int dst[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = tid + r * BLOCK;
        dst[r] = -1;
        if (i < n) {
          const uint32_t nl = newlane[i];
          if (nl != kArrived) {
            dst[r] = (cnt2[nl] & kDirty) ? (int)newidx[i] : (int)start2[nl] + (i - (int)T.lane_start[nl]);
            w[r][0] = __float_as_uint(T.pos[i]); w[r][1] = __float_as_uint(T.speed[i]); w[r][2] = __float_as_uint(T.sf[i]);
            w[r][3] = __float_as_uint(T.tloss[i]); w[r][4] = (uint32_t)T.vid[i]; w[r][5] = T.wr[i]; w[r][6] = T.rc[i];
            w[r][7] = T.meta[i]; w[r][8] = T.ed[i]; w[r][9] = (T.dl[i] & 0xFFFFu) | (nl << 16); w[r][10] = T.aw[i];
          }
        }
      }
      __syncthreads();   // every survivor has been read: the tile may be overwritten in place
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int d = dst[r];
        if (d >= 0) {
          U.pos[d] = __uint_as_float(w[r][0]); U.speed[d] = __uint_as_float(w[r][1]); U.sf[d] = __uint_as_float(w[r][2]);
          U.tloss[d] = __uint_as_float(w[r][3]); U.vid[d] = (int32_t)w[r][4]; U.wr[d] = w[r][5]; U.rc[d] = w[r][6];
          U.meta[d] = w[r][7]; U.ed[d] = w[r][8]; U.dl[d] = w[r][9]; U.aw[d] = w[r][10];
        }
      }
    } else {
    for (int i = tid; i < n; i += BLOCK) {
      uint32_t nl = newlane[i];
      if (nl == kArrived) continue;
      // untouched lane: the segment only shifts; touched lanes were merged above
      int d = (cnt2[nl] & kDirty) ? (int)newidx[i] : (int)start2[nl] + (i - (int)T.lane_start[nl]);
      U.pos[d] = T.pos[i]; U.speed[d] = T.speed[i]; U.sf[d] = T.sf[i]; U.tloss[d] = T.tloss[i];
      U.vid[d] = T.vid[i]; U.wr[d] = T.wr[i]; U.rc[d] = T.rc[i]; U.meta[d] = T.meta[i]; U.ed[d] = T.ed[i];
      U.dl[d] = (T.dl[i] & 0xFFFFu) | (nl << 16); U.aw[d] = T.aw[i];
    }
    }
    for (int j = tid; j < nokc; j += BLOCK) {
      const int o = oklist[j];
      OriginCand c = cand[o];
      if (c.ok_dd < 0) continue;
      int lane = __ldg(sc.origin_lane + o);
      int d = c.rank == kRankBack ? (int)start2[lane + 1] - 1 : (int)start2[lane] + (int)c.rank;
      float dev = sc.speed_dev_override >= 0.0f ? sc.speed_dev_override : VTT(T, c.vt, VT_DEV);
      U.pos[d] = c.pos; U.speed[d] = 0.0f; U.sf[d] = speed_factor(T, c.vid, dev); U.tloss[d] = 0.0f;
      U.vid[d] = c.vid; U.wr[d] = 0u; U.rc[d] = (uint32_t)c.route;
      U.meta[d] = (uint32_t)c.vt | (0xFFu << 16) | (encode_nextlink(sc, lane, choose_link(sc, lane, c.route, 0)) << 24);
      U.ed[d] = 0xFFFEu | ((uint32_t)T.tick << 16);
      U.dl[d] = (uint32_t)(c.ok_dd & 0xFFFF) | ((uint32_t)lane << 16); U.aw[d] = 0u;
    }
  }
  __syncthreads();
  PCLK(PC_S6);

  // ---- S7: swap, bookkeeping ----
  {
    uint32_t* tmp = cur; cur = oth; oth = tmp;
    tile_bind(T, cur, m.vcap);
    uint32_t* occ = (uint32_t*)(smem + m.off_occ);
    for (int dj = tid; dj < ndirty; dj += BLOCK) {
      const int l = dirty[dj];
      const int c = cnt2[l] & ~kDirty;
      cnt2[l] = c; mhead[l] = -1;
      if (c > 0) atomicOr(&occ[l >> 5], 1u << (l & 31)); else atomicAnd(&occ[l >> 5], ~(1u << (l & 31)));
    }
    { uint16_t* t2 = T.lane_start; T.lane_start = start2; start2 = t2; }    // lane offsets ping-pong
    const int n2 = T.lane_start[L];
    if (tid == 0) {
      hdr[H_NVEH] = n2;
      hdr[H_ACTIVE] += n2;
      hdr[H_TICK] += 1;
      if (sc.synthetic) {
        float pend = __int_as_float(hdr[H_F_PENDING]);
        for (int o = 0; o < m.O; ++o) pend += (float)origin_backlog[o];
        hdr[H_F_PENDING] = __float_as_int(pend);
      }
    }
    int anom = 0;
    for (int i = tid; i < n2; i += BLOCK)
      if (i > 0 && (T.dl[i] >> 16) == (T.dl[i - 1] >> 16) && T.pos[i] > T.pos[i - 1]) anom += 1;
    if (anom) atomicAdd(&hdr[H_ANOM], anom);
    T.tick += 1;
  }
  __syncthreads();
  PCLK(PC_S7);
}

// ------------------------------------------------------------------------------------------------
// Signal.observe (traffic_signal.py:189-235) + states.mplight / wave + rewards.* + calc_metrics
template <int BLOCK>
__device__ __forceinline__ void observe_body(const DevSim& D, const SmemLayout& m, unsigned char* smem, unsigned char* vb, Tile& T, int env) {
  const DevScenario& sc = D.sc;
  const int tid = threadIdx.x % BLOCK, lane_id = tid & 31, wid = tid >> 5, nw = BLOCK / 32;
  int32_t* hdr = (int32_t*)(smem + m.off_hdr);
  // an instance that is not stepped by this run (deferred, or taken by the heavy list) must not publish anything: the
  // run that does step it writes its observations, possibly BEFORE this one gets here
  const bool pub = ((const int32_t*)(smem + m.off_misc))[M_BAIL] == 0;
  float* ob = (float*)(m.off_obs == ~(size_t)0 ? vb + m.off_vn : smem + m.off_obs);   // [5][SL]
  const int SL = m.SL, S = m.S;
  const int e = hdr[H_EPOCH];
  const uint32_t eprev = (uint32_t)(e - 1) & 0xFFFFu;
  for (int q = wid; q < SL; q += nw) {     // one warp per inbound lane: segmented shuffle reduction
    int lane = __ldg(sc.sig_lane + q);
    int sg = __ldg(&sc.lane_rec[lane].sig);
    float tdist = __ldg(&sc.lane_rec[lane].tls_dist), llen = __ldg(&sc.lane_rec[lane].len);
    float queue = 0, appr = 0, tw = 0, mw = 0, ss = 0, arrv = 0;
    int a = T.lane_start[lane], b = T.lane_start[lane + 1];
    for (int i = a + lane_id; i < b; i += 32) {
      if (tdist < 0.0f) continue;
      float dist = (llen - T.pos[i]) + tdist;
      if (!(dist <= sc.max_distance)) continue;
      uint32_t w = T.wr[i], mt = T.meta[i], ed = T.ed[i];
      uint32_t rwait = w >> 16, wait = w & 0xFFFFu;
      bool contiguous = ((ed & 0xFFFFu) == eprev) && (((mt >> 16) & 0xFFu) == (uint32_t)sg) && e > 0;
      if (!contiguous) { rwait = 0; arrv += 1.0f; }   // not seen by this signal at its previous observe: an arrival
      if (rwait > 0) rwait += (uint32_t)sc.step_length;
      else if (wait > 0) rwait = wait;
      if (rwait > 0xFFFFu) rwait = 0xFFFFu;
      T.wr[i] = wait | (rwait << 16);
      T.meta[i] = (mt & 0xFF00FFFFu) | ((uint32_t)sg << 16);
      T.ed[i] = (ed & 0xFFFF0000u) | ((uint32_t)e & 0xFFFFu);
      float rw = (float)rwait;
      if (rwait > 0) { tw += rw; queue += 1.0f; mw = fmaxf(mw, rw); } else appr += 1.0f;
      ss += T.speed[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      queue += __shfl_xor_sync(0xffffffffu, queue, o);
      appr += __shfl_xor_sync(0xffffffffu, appr, o);
      tw += __shfl_xor_sync(0xffffffffu, tw, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      arrv += __shfl_xor_sync(0xffffffffu, arrv, o);
      mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    }
    if (lane_id == 0 && !pub) { ob[0 * SL + q] = queue; ob[1 * SL + q] = appr; ob[2 * SL + q] = tw; ob[3 * SL + q] = mw; ob[4 * SL + q] = ss; }
    if (lane_id == 0 && pub) {
      ob[0 * SL + q] = queue; ob[1 * SL + q] = appr; ob[2 * SL + q] = tw; ob[3 * SL + q] = mw; ob[4 * SL + q] = ss;
      size_t g = (size_t)env * SL + q;
      D.lane_queue[g] = queue; D.lane_approach[g] = appr; D.lane_total_wait[g] = tw;
      D.lane_max_wait[g] = mw; D.lane_speed_sum[g] = ss; D.lane_arrivals[g] = arrv;
    }
  }
  __syncthreads();
  if (tid == 0) hdr[H_EPOCH] = e + 1;
  for (int x = tid; x < S * 12 && pub; x += BLOCK) {
    int sg = x / 12, mv = x % 12;
    int q0 = __ldg(sc.sig_lane_off + sg);
    float qsum = 0, wsum = 0;
    for (int j = __ldg(sc.mv_off + x); j < __ldg(sc.mv_off + x + 1); ++j) {
      int q = q0 + __ldg(sc.mv_lane + j);
      qsum += ob[q];
      wsum += ob[q] + ob[SL + q];
    }
    for (int j = __ldg(sc.mvo_off + x); j < __ldg(sc.mvo_off + x + 1); ++j)
      qsum -= ob[__ldg(sc.sig_lane_off + __ldg(sc.mvo_sig + j)) + __ldg(sc.mvo_slot + j)];
    D.mplight[((size_t)env * S + sg) * 13 + 1 + mv] = qsum;
    D.wave[((size_t)env * S + sg) * 12 + mv] = wsum;
    if (D.out_mask & RS_OUT_MPLIGHT_FULL) {   // states.mplight_full (states.py:83-113): speed = the movement's LAST lane's sum
      float wt = 0, spd = 0, ap = 0;
      for (int j = __ldg(sc.mv_off + x); j < __ldg(sc.mv_off + x + 1); ++j) {
        const int q = q0 + __ldg(sc.mv_lane + j);
        wt += ob[2 * SL + q] / 28.0f; spd = ob[4 * SL + q]; ap += ob[SL + q] / 28.0f;
      }
      float* mf = D.mplight_full + ((size_t)env * S + sg) * 49 + 1 + 4 * mv;
      mf[0] = qsum; mf[1] = wt; mf[2] = spd; mf[3] = ap;
    }
  }
  if (D.out_mask & (RS_OUT_DRQ | RS_OUT_DRQ_NORM)) {   // states.drq / drq_norm (states.py:6-59): the one-hot compares the LANE index with the phase
    for (int q = tid; q < SL && pub; q += BLOCK) {
      const int s2 = __ldg(&sc.lane_rec[__ldg(sc.sig_lane + q)].sig);   // rs_create rejects a lane listed by two signals
      const float oh = (q - __ldg(sc.sig_lane_off + s2)) == T.tls_phase[__ldg(sc.sig_tls + s2)] ? 1.0f : 0.0f;
      const float queue = ob[q], appr = ob[SL + q], tw = ob[2 * SL + q], ss = ob[4 * SL + q];
      const size_t g = ((size_t)env * SL + q) * 5;
      if (D.out_mask & RS_OUT_DRQ) { D.drq[g] = oh; D.drq[g + 1] = appr; D.drq[g + 2] = tw; D.drq[g + 3] = queue; D.drq[g + 4] = ss; }
      if (D.out_mask & RS_OUT_DRQ_NORM) {
        D.drq_norm[g] = oh; D.drq_norm[g + 1] = appr / 28.0f; D.drq_norm[g + 2] = tw / 28.0f; D.drq_norm[g + 3] = queue / 28.0f;
        D.drq_norm[g + 4] = ss / 20.0f / 28.0f;
      }
    }
  }
  for (int sg = tid; sg < S && pub; sg += BLOCK) {
    int q0 = __ldg(sc.sig_lane_off + sg), q1 = __ldg(sc.sig_lane_off + sg + 1);
    int ph = T.tls_phase[__ldg(sc.sig_tls + sg)];
    float tw = 0, ql = 0, mq = 0;
    for (int q = q0; q < q1; ++q) { tw += ob[2 * SL + q]; ql += ob[q]; mq = fmaxf(mq, ob[q]); }
    float pr = ql;
    for (int j = __ldg(sc.out_off + sg); j < __ldg(sc.out_off + sg + 1); ++j)
      pr -= ob[__ldg(sc.sig_lane_off + __ldg(sc.out_sig + j)) + __ldg(sc.out_slot + j)];
    size_t g = (size_t)env * S + sg;
    D.phase_obs[g] = ph;
    D.mplight[g * 13] = (float)ph;
    if (D.out_mask & RS_OUT_MPLIGHT_FULL) D.mplight_full[g * 49] = (float)ph;
    D.rew_wait[g] = -tw;
    D.rew_wait_norm[g] = fminf(fmaxf(-tw / 224.0f, -4.0f), 4.0f);
    D.rew_pressure[g] = -pr;
    D.sig_queue_len[g] = (int32_t)ql; D.sig_max_queue[g] = (int32_t)mq;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dev_set_phase(const DevScenario& sc, Tile& T, int sg, int idx) {
  int t = __ldg(sc.sig_tls + sg);
  int p0 = __ldg(sc.tls_phase_off + t), np = __ldg(sc.tls_phase_off + t + 1) - p0;
  if (idx < 0 || idx >= np) return;
  T.tls_phase[t] = idx;
  T.tls_end[t] = T.tick + __ldg(sc.phase_dur + p0 + idx);
  T.tls_state[t] = __ldg(sc.phase_state_off + p0 + idx);
}

// Steps one instance; returns true if the instance outgrew the tile of this run (its HBM state is then untouched).
// `may_defer`: outgrowing the tile defers (to the in-CTA redo or the overflow list) instead of refusing insertions;
// `to_list`: a deferred instance queues itself on the overflow list; `use_tma`: stage with bulk copies + the slot's mbarrier.
template <int BLOCK, int G>
__device__ __forceinline__ bool run_instance(const DevSim& D, const RunArgs& A, const SmemLayout& m,
                                             unsigned char* smem, unsigned char* vb, const int env, const bool real_slot,
                                             uint32_t& tma_parity, const bool may_defer, const bool to_list, const bool use_tma,
                                             const bool skip_heavy) {
  const DevScenario& sc = D.sc;
  const int tid = threadIdx.x % BLOCK;
  uint32_t* cur = (uint32_t*)(vb + m.off_bufA);
  uint32_t* oth = (uint32_t*)(vb + m.off_bufB);
  int32_t* hdr = (int32_t*)(smem + m.off_hdr);
  int32_t* next_phase = (int32_t*)(smem + m.off_next_phase);
  int32_t* origin_cur = (int32_t*)(smem + m.off_origin_cur);
  int32_t* origin_backlog = (int32_t*)(smem + m.off_origin_backlog);
  float* vt = (float*)(smem + m.off_vt);
  Tile T;
  tile_bind(T, cur, m.vcap);
  T.lane_start = (uint16_t*)(smem + m.off_lane_start);
  T.tls_phase = (int32_t*)(smem + m.off_tls_phase);
  T.tls_end = (int32_t*)(smem + m.off_tls_end);
  T.tls_state = (int32_t*)(smem + m.off_tls_state);
  T.vt = vt;
  T.occ = (const uint32_t*)(smem + m.off_occ);
  const uint64_t env_id = (uint64_t)(D.first_env_id + env);
  T.env_lo = (uint32_t)env_id; T.env_hi = (uint32_t)(env_id >> 32);
  T.seed_lo = (uint32_t)D.seed; T.seed_hi = (uint32_t)(D.seed >> 32);

  // ---- stage the instance tile: HBM -> shared (128-bit coalesced loads) ----
  if (tid < kHdrInts) hdr[tid] = D.hdr[(size_t)env * kHdrInts + tid];
  for (int i = tid; i < m.n_tls; i += BLOCK) {
    T.tls_phase[i] = D.tls_phase[(size_t)env * m.n_tls + i];
    T.tls_end[i] = D.tls_end[(size_t)env * m.n_tls + i];
  }
  for (int i = tid; i < m.S; i += BLOCK) next_phase[i] = D.next_phase[(size_t)env * m.S + i];
  for (int i = tid; i < m.O; i += BLOCK) {
    const int oc = D.origin_cur[(size_t)env * m.O + i];
    origin_cur[i] = oc;
    if (sc.synthetic) origin_backlog[i] = D.origin_backlog[(size_t)env * m.O + i];
    else {   // departure time of the next trip of this origin (see S3a)
      const int ci = __ldg(sc.origin_off + i) + oc;
      origin_backlog[i] = __float_as_int(ci < __ldg(sc.origin_off + i + 1) ? __ldg(sc.trip_depart + ci) : 3.0e38f);
    }
  }
  for (int i = tid; i < m.n_vt * 8; i += BLOCK) vt[i] = __ldg(sc.vtype + i);
  __syncthreads();
  int32_t* misc0 = (int32_t*)(smem + m.off_misc);
  if (tid == 0) {   // already larger than this launch's tile: defer at once and step an empty tile (barriers stay in lock-step)
    // skip_heavy (2): the instance is on this launch's heavy list -- it was above the heavy threshold when the previous
    // launch wrote it back -- and the CTA that took it there either has stepped it already (launch stamp) or still is
    const int stamp = D.heavy_count ? D.heavy_count[3] + 1 : 0;
    const int big = (!real_slot || (skip_heavy && (hdr[H_DONE] == stamp || hdr[H_NVEH] > D.heavy_thr))) ? 2
                  : ((may_defer && hdr[H_NVEH] > m.vcap) ? 1 : 0);
    misc0[M_BAIL] = big; misc0[M_MAYDEFER] = may_defer ? 1 : 0;
    if (big) hdr[H_NVEH] = 0;
  }
  __syncthreads();
  const int n0 = hdr[H_NVEH];
  {
    const uint32_t* g = D.veh + (size_t)env * kVehWords * sc.vcap;
    const int n4 = (n0 + 3) >> 2;
    if (use_tma) {   // TMA: 1-D bulk copies (one per word array) tracked by the slot's mbarrier
      if (n4 > 0) {
        uint64_t* bar = (uint64_t*)(smem + m.off_mbar);
        if (tid == 0) {
          mbar_expect_tx(bar, (uint32_t)(kVehWords * n4 * 16));
          for (int w = 0; w < kVehWords; ++w)
            tma_load_1d(cur + (size_t)w * m.vcap, g + (size_t)w * sc.vcap, (uint32_t)(n4 * 16), bar);
        }
        mbar_wait(bar, tma_parity);
        tma_parity ^= 1u;
      }
    } else {
      for (int w = 0; w < kVehWords; ++w) {
        const uint4* src = (const uint4*)(g + (size_t)w * sc.vcap);
        uint4* dst = (uint4*)(cur + (size_t)w * m.vcap);
        for (int i = tid; i < n4; i += BLOCK) dst[i] = __ldcs(src + i);
      }
    }
  }
  T.tick = hdr[H_TICK];
  __syncthreads();
  // lane offsets from the per-vehicle lane ids (vehicles are stored lane-major)
  for (int l = tid; l <= m.L; l += BLOCK) {
    int lo = 0, hi = n0;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((int)(T.dl[mid] >> 16) < l) lo = mid + 1; else hi = mid; }
    T.lane_start[l] = (uint16_t)lo;
  }
  __syncthreads();

  // per-lane counters stay current from tick to tick; only the lanes a tick touches are revisited
  uint16_t* start2 = (uint16_t*)(smem + m.off_start2);
  {
    int32_t* cnt2 = (int32_t*)(smem + m.off_cnt2);
    int32_t* mhead = (int32_t*)(smem + m.off_mhead);
    for (int l = tid; l < m.L; l += BLOCK) { cnt2[l] = lane_count(T, l); mhead[l] = -1; }
    // lane-occupancy bits (kept current per tick through the dirty-lane list): one ballot per 32 lanes
    uint32_t* occ = (uint32_t*)(smem + m.off_occ);
    for (int base = 0; base < m.L + 64; base += BLOCK) {
      const int l = base + tid;
      const uint32_t word = __ballot_sync(0xFFFFFFFFu, l < m.L && lane_count(T, l) > 0);
      if ((tid & 31) == 0 && (l >> 5) < (m.L + 31) / 32 + 2) occ[l >> 5] = word;
    }
  }
  __syncthreads();

  // ---- MultiSignal.step schedule (multi_signal.py:164-197) ----
  if (A.do_prep) {
    for (int sg = tid; sg < m.S; sg += BLOCK) {   // Signal.prep_phase (traffic_signal.py:176-184)
      int act = A.actions[(size_t)env * m.S + sg];
      int cp = T.tls_phase[__ldg(sc.sig_tls + sg)];
      if (cp == act) next_phase[sg] = cp;
      else {
        next_phase[sg] = act;
        int ng = __ldg(sc.sig_n_green + sg);
        if (cp >= 0 && cp < ng && act >= 0 && act < ng) {
          int y = __ldg(sc.yellow_idx + __ldg(sc.sig_yellow_off + sg) + cp * ng + act);
          if (y >= 0) dev_set_phase(sc, T, sg, y);
        }
      }
    }
    __syncthreads();
  }
  const int n_ticks = A.ticks_a + A.ticks_b;
  PCLK(PC_STAGE);
#pragma unroll 1
  for (int k = 0; k <= n_ticks; ++k) {
    if (k == A.ticks_a && A.do_set) {   // Signal.set_phase after the yellow interval
      for (int sg = tid; sg < m.S; sg += BLOCK) dev_set_phase(sc, T, sg, next_phase[sg]);
      __syncthreads();
    }
    if (k == n_ticks) break;
    tick_body<BLOCK, G>(D, m, smem, vb, T, cur, oth, start2, env);
  }
  if (A.do_observe) observe_body<BLOCK>(D, m, smem, vb, T, env);
  PCLK(PC_OBS);

  // ---- write the tile back; a deferred instance leaves its HBM state untouched and queues itself for the overflow
  //      pass (same barriers either way: the instances of a CTA run in lock-step) ----
  const bool bail = misc0[M_BAIL] != 0;
  if (misc0[M_BAIL] == 1 && to_list && tid == 0 && real_slot) D.overflow_list[atomicAdd(D.overflow_count, 1)] = env;
  if (!bail && D.heavy_count && tid == 0 && real_slot && hdr[H_NVEH] > D.heavy_thr)   // heavy for the NEXT launch
    { const int hw = D.heavy_count[2] ^ 1; D.heavy_list[hw][atomicAdd(D.heavy_count + hw, 1)] = env; }
  {
    const int n1 = hdr[H_NVEH];
    uint32_t* g = D.veh + (size_t)env * kVehWords * sc.vcap;
    const int n4 = (n1 + 3) >> 2;
    if (use_tma) {   // shared -> global bulk stores; the tile may be reused once they have been READ
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0 && n4 > 0 && !bail) {
        for (int w = 0; w < kVehWords; ++w)
          tma_store_1d(g + (size_t)w * sc.vcap, cur + (size_t)w * m.vcap, (uint32_t)(n4 * 16));
        tma_store_commit_and_wait_read();
      }
    } else if (!bail) {
      for (int w = 0; w < kVehWords; ++w) {
        uint4* dst = (uint4*)(g + (size_t)w * sc.vcap);
        const uint4* src = (const uint4*)(cur + (size_t)w * m.vcap);
        for (int i = tid; i < n4; i += BLOCK) __stcs(dst + i, src[i]);
      }
    }
  }
  if (!bail) {
    if (tid == 0 && D.heavy_count) hdr[H_DONE] = D.heavy_count[3] + 1;
    __syncwarp();
    if (tid < kHdrInts) D.hdr[(size_t)env * kHdrInts + tid] = hdr[tid];
    for (int i = tid; i < m.n_tls; i += BLOCK) {
      D.tls_phase[(size_t)env * m.n_tls + i] = T.tls_phase[i];
      D.tls_end[(size_t)env * m.n_tls + i] = T.tls_end[i];
    }
    for (int i = tid; i < m.S; i += BLOCK) D.next_phase[(size_t)env * m.S + i] = next_phase[i];
    for (int i = tid; i < m.O; i += BLOCK) {
      D.origin_cur[(size_t)env * m.O + i] = origin_cur[i];
      if (sc.synthetic) D.origin_backlog[(size_t)env * m.O + i] = origin_backlog[i];
    }
  }
  PCLK(PC_WRITE);
  return misc0[M_BAIL] == 1;
}

// Persistent launch: the grid holds as many CTAs as are resident at once (148 SMs x CTAs/SM); each CTA
// steps G instances in lock-step (TPI threads each: one instruction stream and one set of barriers serve
// G instances, which keeps the instruction cache and the warp slots busy) and pulls the next G from a
// global counter until all N are done -- no partial last wave, uneven instances balance out.
template <int TPI, int G, int MINB>
__global__ void __launch_bounds__(TPI * G, MINB) k_run(const __grid_constant__ DevSim D, const __grid_constant__ RunArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout m = make_layout(D.sc);
  unsigned char* my = smem + (size_t)(threadIdx.x / TPI) * m.total;
  // per-vehicle region: the instance slot's shared memory, or (tile_gmem) this CTA's slot of the global workspace
  unsigned char* vb = m.gmem ? D.workspace + ((size_t)blockIdx.x * G + threadIdx.x / TPI) * m.veh_total : my;
  __shared__ int s_env;
  uint32_t tma_parity = 0;
#if RS_PHASE_CLOCKS
  if (threadIdx.x == 0) { for (int i = 0; i < 24; ++i) s_pclk[i] = 0; s_pclk_last = clock64(); }
#endif
  if (D.use_tma) {
    if (threadIdx.x % TPI == 0) mbar_init((uint64_t*)(my + m.off_mbar), 1);
    __syncthreads();
  }
  const bool has_next = D.overflow_count != nullptr && !D.from_list;      // somebody can take what outgrows this tile
  const bool redo = G > 1 && D.sc.redo_cap > 0;
  const bool tma = D.use_tma && !m.gmem;
  if constexpr (G > 1) {
    // ---- the heavy instances of the previous launch first: each by the whole CTA on the redo tile ----
    if (redo && D.heavy_count && D.heavy_count[D.heavy_count[2]] > 0) {
      const int hc = D.heavy_count[2];
      const int nh = D.heavy_count[hc];
      const SmemLayout mb = make_layout_ex(D.sc, D.sc.redo_cap, D.sc.redo_single, 0);
      bool any = false;
      uint32_t unused_parity = 0;
      for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_env = atomicAdd(D.heavy_taken, 1);
        __syncthreads();
        const int idx = s_env;
        if (idx >= nh) break;
        if (!any && tma) { if (threadIdx.x % TPI == 0) mbar_inval((uint64_t*)(my + m.off_mbar)); __syncthreads(); }
        any = true;
        if (threadIdx.x == 0 && D.redo_count) atomicAdd(D.redo_count, 1);
        run_instance<TPI * G, 1>(D, A, mb, smem, smem, D.heavy_list[hc][idx], true, unused_parity, has_next, has_next, false, false);
      }
      if (any && tma) {
        __syncthreads();
        if (threadIdx.x % TPI == 0) mbar_init((uint64_t*)(my + m.off_mbar), 1);
      }
      __syncthreads();
    }
  }
#pragma unroll 1
  for (;;) {
    const int n_work = D.from_list ? *D.overflow_count : D.n_env;
    if (D.persistent) {
      if (threadIdx.x == 0) s_env = atomicAdd(D.work_counter, G);
      __syncthreads();
    }
    PCLK(PC_SCHED);
    const int env0 = D.persistent ? s_env : (int)blockIdx.x * G;
    if (env0 >= n_work) break;
    // an idle slot (past the end of the batch) steps an empty tile and stores nothing
    const int slot_i = (int)(threadIdx.x / TPI);
    const int slot = env0 + slot_i;
    const bool live = slot < n_work;
    const int item = min(slot, n_work - 1);
    const int env = D.from_list ? D.overflow_list[item] : item;
    const bool bail = run_instance<TPI, G>(D, A, m, my, vb, env, live, tma_parity, has_next || redo, has_next && !redo, tma,
                                           redo && D.heavy_count != nullptr);
    if constexpr (G > 1) {
      if (redo) {   // instances of this group that outgrew their slot: the whole CTA steps them again, one at a time
        __shared__ int s_redo[G];
        __shared__ int s_nredo;
        __syncthreads();
        if (threadIdx.x == 0) s_nredo = 0;
        __syncthreads();
        if (bail && live && threadIdx.x % TPI == 0) s_redo[atomicAdd(&s_nredo, 1)] = env;
        __syncthreads();
        const int nr = s_nredo;
        if (nr > 0) {   // (s_redo is static shared memory: the redo tile only overwrites the dynamic part)
          if (threadIdx.x == 0 && D.redo_count) atomicAdd(D.redo_count, nr);
          if (tma) { if (threadIdx.x % TPI == 0) mbar_inval((uint64_t*)(my + m.off_mbar)); }
          __syncthreads();
          const SmemLayout mb = make_layout_ex(D.sc, D.sc.redo_cap, D.sc.redo_single, 0);
          uint32_t unused_parity = 0;
          for (int r = 0; r < nr; ++r) {
            run_instance<TPI * G, 1>(D, A, mb, smem, smem, s_redo[r], true, unused_parity, has_next, has_next, false, false);
            __syncthreads();
          }
          if (tma) {
            if (threadIdx.x % TPI == 0) mbar_init((uint64_t*)(my + m.off_mbar), 1);
            tma_parity = 0;
          }
          __syncthreads();
        }
      }
    }
    if (!D.persistent) break;
    __syncthreads();
  }
#if RS_PHASE_CLOCKS
  if (threadIdx.x == 0 && D.phase_clocks)
    for (int i = 0; i < PC_N; ++i) atomicAdd(D.phase_clocks + i, (unsigned long long)s_pclk[i]);
#endif
}

// ------------------------------------------------------------------------------------------------
__global__ void k_reset(DevSim D) {
  const int env = blockIdx.x;
  const DevScenario& sc = D.sc;
  for (int i = threadIdx.x; i < kHdrInts; i += blockDim.x) D.hdr[(size_t)env * kHdrInts + i] = 0;
  for (int i = threadIdx.x; i < sc.n_tls; i += blockDim.x) {
    D.tls_phase[(size_t)env * sc.n_tls + i] = sc.tls_init_phase[i];
    D.tls_end[(size_t)env * sc.n_tls + i] = sc.tls_init_left[i];
  }
  for (int i = threadIdx.x; i < sc.n_signals; i += blockDim.x) D.next_phase[(size_t)env * sc.n_signals + i] = 0;
  for (int i = threadIdx.x; i < sc.n_origins; i += blockDim.x) {
    D.origin_cur[(size_t)env * sc.n_origins + i] = 0;
    D.origin_backlog[(size_t)env * sc.n_origins + i] = 0;
  }
  if (D.trip_rec)
    for (int i = threadIdx.x; i < sc.n_trips; i += blockDim.x) { D.trip_rec[(size_t)env * sc.n_trips + i] = make_int4(-1, 0, 0, 0); D.trip_wait[(size_t)env * sc.n_trips + i] = 0.0f; }
}

__global__ void k_set_phase(DevSim D, const int32_t* phase, const uint8_t* mask) {
  const DevScenario& sc = D.sc;
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= D.n_env * sc.n_signals) return;
  if (mask && !mask[x]) return;
  int env = x / sc.n_signals, sg = x % sc.n_signals;
  int t = sc.sig_tls[sg];
  int p0 = sc.tls_phase_off[t], np = sc.tls_phase_off[t + 1] - p0;
  int idx = phase[x];
  if (idx < 0 || idx >= np) return;
  D.tls_phase[(size_t)env * sc.n_tls + t] = idx;
  D.tls_end[(size_t)env * sc.n_tls + t] = D.hdr[(size_t)env * kHdrInts + H_TICK] + sc.phase_dur[p0 + idx];
}

// per-instance episode statistics: sequential loops in storage order (matches the oracle's order)
__global__ void k_stats(DevSim D, RsStats* out) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= D.n_env) return;
  const DevScenario& sc = D.sc;
  const int32_t* h = D.hdr + (size_t)env * kHdrInts;
  RsStats st;
  st.tick = h[H_TICK]; st.n_active = h[H_NVEH]; st.n_inserted = h[H_NINS]; st.n_arrived = h[H_NARR];
  st.anomalies = h[H_ANOM]; st.sum_active_ticks = h[H_ACTIVE];
  st.sum_delay_arrived = __int_as_float(h[H_F_DELAY_ARR]);
  st.sum_duration_arrived = __int_as_float(h[H_F_DUR_ARR]);
  st.sum_wait_arrived = __int_as_float(h[H_F_WAIT_ARR]);
  st.n_cap_refused = h[H_NREF];
  const uint32_t* g = D.veh + (size_t)env * kVehWords * sc.vcap;
  const float* tloss = (const float*)(g + 3 * (size_t)sc.vcap);
  const uint32_t* dl = g + 9 * (size_t)sc.vcap;
  float run = 0;
  for (int i = 0; i < st.n_active; ++i) run += tloss[i] + (float)(dl[i] & 0xFFFFu);
  st.sum_delay_running = run;
  int backlog = 0;
  if (!sc.synthetic) {
    float pend = 0;
    for (int o = 0; o < sc.n_origins; ++o)
      for (int c = sc.origin_off[o] + D.origin_cur[(size_t)env * sc.n_origins + o]; c < sc.origin_off[o + 1]; ++c)
        if (sc.trip_depart[c] <= (float)st.tick) { pend += (float)st.tick - sc.trip_depart[c]; backlog++; }
    st.sum_delay_pending = pend;
  } else {
    st.sum_delay_pending = __int_as_float(h[H_F_PENDING]);
    for (int o = 0; o < sc.n_origins; ++o) backlog += D.origin_backlog[(size_t)env * sc.n_origins + o];
  }
  st.n_backlog = backlog;
  out[env] = st;
}

// after every env-step launch: the list the launch filled becomes the one the next launch reads
__global__ void k_heavy_flip(int32_t* heavy_count) {
  const int cur = heavy_count[2] ^ 1;
  heavy_count[2] = cur;
  heavy_count[cur ^ 1] = 0;
  heavy_count[3] += 1;
}

// Batched WaveAgent.act (agents/maxwave.py:18-38): first maximum, in the reference's evaluation order
// (the iteration order of valid_acts[signal]), of obs[p0] + obs[p1] over the valid phase pairs.
// obs = states.mplight[1:] (MAXPRESSURE, agents/maxpressure.py:13-18) or states.wave (MAXWAVE).
// order: [S][n_pairs][2] = (pair index, action) in evaluation order, pair index -1 terminates.
__global__ void k_policy(DevSim D, const int32_t* pairs, int n_pairs, const int32_t* order, int use_wave,
                         int32_t* actions) {
  const DevScenario& sc = D.sc;
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= D.n_env * sc.n_signals) return;
  int sg = x % sc.n_signals;
  const float* ob = use_wave ? D.wave + (size_t)x * 12 : D.mplight + (size_t)x * 13 + 1;
  float best = 0.0f; int bi = -1;
  for (int k = 0; k < n_pairs; ++k) {
    int p = order[(sg * n_pairs + k) * 2];
    if (p < 0) break;
    float pr = ob[pairs[2 * p]] + ob[pairs[2 * p + 1]];
    if (bi < 0 || pr > best) { best = pr; bi = order[(sg * n_pairs + k) * 2 + 1]; }
  }
  actions[x] = bi < 0 ? 0 : bi;
}

}  // namespace rs

// ================================================================================================
// C-ABI
// ================================================================================================
using namespace rs;

struct RsSim {
  DevSim d;
  SmemLayout layout;
  int device;
  int block;   // threads per instance
  int group;   // instances per CTA
  int minb;
  int carveout;
  int smem_extra;   // RESCO_B200_SMEM_EXTRA: unused dynamic shared memory per CTA (experiments on the L1 / shared split)
  int n_sm;
  int resident_ctas;
  // overflow pass (see rs_create): same kernel, whole store in the global workspace, instances from overflow_list
  bool two_pass; SmemLayout layout2; int block2, group2, minb2, resident_ctas2; unsigned char* workspace2;
  int32_t* counters;   // [0] work counter of the fast pass, [1] of the overflow pass, [2] instances deferred to it, [3] redone
                       // in-CTA, [4] work counter over the heavy list; zeroed before every launch
  std::vector<void*> allocs;
  int64_t launches;
  cudaEvent_t ev0, ev1, ev_done;
  bool timed;
  bool pending; float* pend_obs; float* pend_rew;   // rs_env_step_host_async -> rs_wait
  // host staging for rs_env_step_host
  int32_t* h_act_pinned; float* h_obs_pinned; float* h_rew_pinned;
  int32_t* d_actions;
  RsStats* d_stats;
  int32_t *d_pairs, *d_valid; int n_pairs_alloc;
  int host_obs_kind; size_t host_obs_floats;   // rs_set_host_obs
  int use_wave_tables;                          // tables of the last rs_policy_maxpressure call
  float* d_frap; int32_t* d_frap_pairs; uint8_t* d_frap_comp; int32_t* d_frap_order; int frap_pairs;   // rs_frap_load
  cudaGraphExec_t graph; int graph_policy; uint64_t graph_seed;   // rs_env_step_policy
};

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) \
  return fail(RS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } while (0)

extern "C" const char* rs_last_error(void) { return g_err.c_str(); }
extern "C" int rs_abi_version(void) { return RS_ABI_VERSION; }

template <typename T>
static int dev_dup(RsSim* s, const T*& field, size_t count) {
  size_t bytes = sizeof(T) * (count ? count : 1);
  void* p = nullptr;
  CK(cudaMalloc(&p, bytes));
  s->allocs.push_back(p);
  if (field && count) CK(cudaMemcpy(p, field, sizeof(T) * count, cudaMemcpyHostToDevice));
  else CK(cudaMemset(p, 0, bytes));
  field = (const T*)p;
  return 0;
}
template <typename T>
static int dev_alloc(RsSim* s, T*& out, size_t count) {
  void* p = nullptr;
  size_t bytes = sizeof(T) * (count ? count : 1);
  CK(cudaMalloc(&p, bytes));
  CK(cudaMemset(p, 0, bytes));
  s->allocs.push_back(p);
  out = (T*)p;
  return 0;
}
#define TRY(x) do { int _r = (x); if (_r) return _r; } while (0)

template <int TPI, int G, int MINB>
static int launch_run(RsSim* s, const DevSim& d, size_t smem_per_instance, int resident, int n_work, const RunArgs& a, cudaStream_t st) {
  int grid = (n_work + G - 1) / G;
  if (d.persistent && resident < grid) grid = resident;
  k_run<TPI, G, MINB><<<grid, TPI * G, smem_per_instance * G + s->smem_extra, st>>>(d, a);
  s->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// (threads per instance, instances per CTA, min CTAs/SM for __launch_bounds__: 1 = registers uncapped,
//  1024/(TPI*G) = 64 registers per thread)
//  The shapes rs_create can choose (fit_group: 8, 7, 6, 5 instances of 64 threads, 4 of 128, 2 or 1 of 512); other
//  combinations were measured in round 1 and are no longer compiled.
#define RS_VARIANTS(X) X(64, 5, 1) X(64, 6, 1) X(64, 7, 1) X(64, 8, 1) X(128, 4, 1) X(512, 1, 1) X(512, 2, 1) X(512, 1, 2)

static int launch_variant(RsSim* s, int block, int group, int minb, const DevSim& d, size_t smem_per_instance, int resident,
                          int n_work, const RunArgs& a, cudaStream_t st) {
#define X(B, G, M) if (block == B && group == G && minb == M) return launch_run<B, G, M>(s, d, smem_per_instance, resident, n_work, a, st);
  RS_VARIANTS(X)
#undef X
  return fail(RS_ERR_INVALID, "unsupported RESCO_B200_BLOCK / RESCO_B200_GROUP / RESCO_B200_REGCAP combination");
}

// the overflow pass's view of the sim: same state, the whole store as its tile (global workspace), work from the list
static DevSim overflow_view(const RsSim* s) {
  DevSim d2 = s->d;
  d2.sc.tile_cap = d2.sc.vcap; d2.sc.tile_single = 0; d2.sc.tile_gmem = 1;
  d2.workspace = s->workspace2;
  d2.work_counter = s->counters + 1;
  d2.overflow_count = s->counters + 2; d2.from_list = 1;   // from_list: the count is read, nothing is deferred again
  d2.persistent = 1;
  return d2;
}

static int run(RsSim* s, const RunArgs& a, cudaStream_t st) {
  if (s->d.persistent) CK(cudaMemsetAsync(s->counters, 0, 8 * sizeof(int32_t), st));
  TRY(launch_variant(s, s->block, s->group, s->minb, s->d, s->layout.total, s->resident_ctas, s->d.n_env, a, st));
  if (s->two_pass)
    TRY(launch_variant(s, s->block2, s->group2, s->minb2, overflow_view(s), s->layout2.total, s->resident_ctas2, s->d.n_env, a, st));
  if (s->d.heavy_count) { k_heavy_flip<<<1, 1, 0, st>>>(s->d.heavy_count); s->launches += 1; CK(cudaGetLastError()); }
  return 0;
}

static int configure(RsSim* s) {
  // shared memory per CTA of each compiled shape = the larger of the passes that use it
  auto bytes_of = [&](int B, int G, int M) {
    size_t b = 0;
    if (s->block == B && s->group == G && s->minb == M) b = s->layout.total * G + s->smem_extra;
    if (s->two_pass && s->block2 == B && s->group2 == G && s->minb2 == M) { const size_t b2 = s->layout2.total * G + s->smem_extra; b = b2 > b ? b2 : b; }
    return (int)b;
  };
  bool found = false;
#define X(B, G, M) { const int bytes = bytes_of(B, G, M); if (bytes > 0) { \
    CK(cudaFuncSetAttribute(k_run<B, G, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
    if (s->carveout >= 0) CK(cudaFuncSetAttribute(k_run<B, G, M>, cudaFuncAttributePreferredSharedMemoryCarveout, s->carveout)); \
    if (s->block == B && s->group == G && s->minb == M) { \
      int per_sm = 0; \
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_run<B, G, M>, B * G, s->layout.total * G + s->smem_extra)); \
      s->resident_ctas = (per_sm > 0 ? per_sm : 1) * s->n_sm; found = true; } \
    if (s->two_pass && s->block2 == B && s->group2 == G && s->minb2 == M) { \
      int per_sm = 0; \
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_run<B, G, M>, B * G, s->layout2.total * G + s->smem_extra)); \
      s->resident_ctas2 = (per_sm > 0 ? per_sm : 1) * s->n_sm; } } }
  RS_VARIANTS(X)
#undef X
  if (!found) return fail(RS_ERR_INVALID, "unsupported RESCO_B200_BLOCK / RESCO_B200_GROUP / RESCO_B200_REGCAP combination");
  return 0;
}

extern "C" int rs_create(const RsScenario* sc, int32_t n_env, int32_t device, uint64_t seed, RsSim** out) {
  if (!sc || !out || n_env <= 0) return fail(RS_ERR_INVALID, "rs_create: bad arguments");
  if (sc->abi_version != RS_ABI_VERSION) return fail(RS_ERR_INVALID, "rs_create: ABI version mismatch");
  if (sc->vcap <= 0 || sc->vcap % 4 || sc->vcap > 65528) return fail(RS_ERR_INVALID, "rs_create: vcap must be a positive multiple of 4 (< 65528)");
  if (sc->n_lanes >= 65535 || sc->n_signals > 254 || sc->n_routes > 65535 || sc->n_vtypes > 255)
    return fail(RS_ERR_INVALID, "rs_create: scenario exceeds the packed-field ranges");
  {   // the per-lane sweep finds a row's signal through the lane (LaneRec::sig): one signal per inbound lane
    std::vector<char> seen((size_t)sc->n_lanes, 0);
    for (int q = 0; q < sc->n_sig_lanes; ++q) {
      const int l = sc->sig_lane[q];
      if (l < 0 || l >= sc->n_lanes || seen[l]) return fail(RS_ERR_INVALID, "rs_create: an inbound lane is listed twice in sig_lane");
      seen[l] = 1;
    }
  }
  {   // one origin per lane: a tick inserts at most one vehicle into a lane
    std::vector<char> seen((size_t)sc->n_lanes, 0);
    for (int o = 0; o < sc->n_origins; ++o) {
      const int l = sc->origin_lane[o];
      if (l < 0 || l >= sc->n_lanes || seen[l]) return fail(RS_ERR_INVALID, "rs_create: a lane is listed twice in origin_lane");
      seen[l] = 1;
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(RS_ERR_NODEVICE, "rs_create: no CUDA device (this backend has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(RS_ERR_INVALID, "rs_create: bad device index");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(RS_ERR_NODEVICE, "rs_create: built for sm_100a (Blackwell) only");
  RsSim* s = new RsSim();
  s->device = device; s->launches = 0; s->timed = false; s->n_pairs_alloc = 0;
  s->pending = false; s->pend_obs = nullptr; s->pend_rew = nullptr; s->ev_done = nullptr;
  s->d_frap = nullptr; s->frap_pairs = 0; s->graph = nullptr; s->graph_policy = 0; s->graph_seed = 0; s->use_wave_tables = 0;
  static_cast<RsScenario&>(s->d.sc) = *sc; s->d.n_env = n_env; s->d.seed = seed; s->d.first_env_id = 0;
  DevScenario& d = s->d.sc;
  const int L = sc->n_lanes, K = sc->n_links, S = sc->n_signals;
  {   // packed device tables (LaneRec / LinkRec / FoeRec / route_step_link), built from the caller's host arrays
    std::vector<LaneRec> lr((size_t)L + 1);
    std::vector<LinkRec> kr((size_t)K + 1);
    std::vector<FoeRec> fr((size_t)sc->n_foes + 1);
    std::vector<uint8_t> rsl((size_t)(sc->n_route_steps > 0 ? sc->n_route_steps : 1) * 8, 0xFDu);
    memset(lr.data(), 0, lr.size() * sizeof(LaneRec)); memset(kr.data(), 0, kr.size() * sizeof(LinkRec));
    memset(fr.data(), 0, fr.size() * sizeof(FoeRec));
    for (int l = 0; l <= L; ++l) {
      LaneRec& r = lr[l];
      r.link_off = sc->lane_link_off[l];
      if (l == L) break;
      r.len = sc->lane_len[l]; r.vmax = sc->lane_vmax[l]; r.internal = sc->lane_internal[l];
      r.index = sc->lane_index[l]; r.left = sc->lane_left[l]; r.right = sc->lane_right[l]; r.perm = sc->lane_perm[l];
      r.tls_dist = sc->lane_tls_dist[l]; r.sig = sc->lane_sig[l]; r.sig_slot = sc->lane_sig_slot[l];
    }
    for (int k = 0; k <= K; ++k) {
      LinkRec& r = kr[k];
      r.foe_off = sc->link_foe_off[k];
      if (k == K) break;
      r.from = sc->link_from[k]; r.to = sc->link_to[k]; r.via = sc->link_via[k]; r.to_edge = sc->link_to_edge[k];
      r.tls = sc->link_tls[k]; r.tlidx = sc->link_tlidx[k]; r.state = sc->link_state[k]; r.cont = sc->link_cont[k];
      r.via_len = sc->link_via_len[k]; r.last_int = sc->link_last_int[k]; r.parent = sc->link_parent[k];
      r.nxt = r.via >= 0 ? r.via : r.to;
      r.nxt_len = sc->lane_len[r.nxt]; r.nxt_vmax = sc->lane_vmax[r.nxt];
      r.nxt_internal = sc->lane_internal[r.nxt]; r.from_internal = sc->lane_internal[r.from];
      r.nxt_link = !r.nxt_internal ? -3 : (sc->lane_link_off[r.nxt] < sc->lane_link_off[r.nxt + 1] ? sc->lane_link_off[r.nxt] : -2);
      r.yield_parent = -1; r.yield_cross = 0.0f;
      if (r.from_internal) {
        const int p = sc->link_parent[k];
        if (p >= 0 && sc->link_cont[p] && sc->link_via[p] == r.from) {
          r.yield_parent = p;
          r.yield_cross = sc->link_via_len[p] - sc->lane_len[r.from];
        }
      }
      r.foe_end = sc->link_foe_off[k + 1];
      int lo = 0x7FFFFFFF, hi = -1;
      for (int i = sc->link_foe_off[k]; i < sc->link_foe_off[k + 1]; ++i) {
        const int li = sc->link_last_int[sc->foe_link[i]];
        if (li >= 0) { lo = li < lo ? li : lo; hi = li > hi ? li : hi; }
      }
      r.occ_word = hi < 0 ? 0 : lo >> 5; r.occ_lo = 0u; r.occ_hi = 0u;
      if (hi >= 0 && (hi >> 5) - (lo >> 5) > 1) r.occ_word = -1;
      else
        for (int i = sc->link_foe_off[k]; i < sc->link_foe_off[k + 1]; ++i) {
          const int li = sc->link_last_int[sc->foe_link[i]];
          if (li < 0) continue;
          if ((li >> 5) == r.occ_word) r.occ_lo |= 1u << (li & 31); else r.occ_hi |= 1u << (li & 31);
        }
    }
    for (int i = 0; i < sc->n_foes; ++i) {
      FoeRec& r = fr[i];
      const int f = sc->foe_link[i];
      r.link = f; r.flags = (sc->foe_flags[i] & 7) | (sc->link_cont[f] ? 8 : 0);
      r.last_int = sc->link_last_int[f]; r.from = sc->link_from[f];
      r.slot = sc->link_cont[f] ? sc->link_via[f] : -1; r.via_len = sc->link_via_len[f];
      r.len_from = sc->lane_len[r.from]; r.len_slot = r.slot >= 0 ? sc->lane_len[r.slot] : 0.0f;
    }
    fr[sc->n_foes].last_int = -1; fr[sc->n_foes].slot = -1;
    // choose_link tabulated per (route step, lane index): prefer a target lane that is "best", then "ok", then any
    for (int r = 0; r < sc->n_routes; ++r) {
      const int ro = sc->route_off[r], rn = sc->route_off[r + 1] - ro;
      for (int c = 0; c < rn; ++c) {
        const int e = sc->route_edge[ro + c];
        for (int j = 0; j < sc->edge_nlanes[e] && j < 8; ++j) {
          const int lane = sc->edge_lane0[e] + j;
          if (sc->lane_index[lane] != j || sc->lane_internal[lane]) {
            rs_destroy(s);
            return fail(RS_ERR_INVALID, "rs_create: lanes of a route edge are not contiguous by index");
          }
          int code;
          if (c + 1 >= rn) code = 0xFE;
          else {
            const int ne = sc->route_edge[ro + c + 1], mask = sc->route_mask[ro + c + 1];
            int best = -2, rank = 0;
            for (int k = sc->lane_link_off[lane]; k < sc->lane_link_off[lane + 1]; ++k) {
              if (sc->link_to_edge[k] != ne) continue;
              const int ti = sc->lane_index[sc->link_to[k]];
              const int q = ((mask >> (8 + ti)) & 1) ? 3 : (((mask >> ti) & 1) ? 2 : 1);
              if (q > rank) { rank = q; best = k; }
            }
            code = best < 0 ? 0xFD : best - sc->lane_link_off[lane];
            if (best >= 0 && code >= 0xFD) { rs_destroy(s); return fail(RS_ERR_INVALID, "rs_create: more than 252 links on one lane"); }
          }
          rsl[(size_t)(ro + c) * 8 + j] = (uint8_t)code;
        }
      }
    }
    const LaneRec* lp = lr.data(); const LinkRec* kp = kr.data(); const FoeRec* fp = fr.data(); const uint8_t* rp = rsl.data();
    TRY(dev_dup(s, lp, lr.size())); TRY(dev_dup(s, kp, kr.size())); TRY(dev_dup(s, fp, fr.size())); TRY(dev_dup(s, rp, rsl.size()));
    d.lane_rec = lp; d.link_rec = kp; d.foe_rec = fp; d.route_step_link = rp;
  }
  TRY(dev_dup(s, d.lane_len, L)); TRY(dev_dup(s, d.lane_vmax, L)); TRY(dev_dup(s, d.lane_edge, L));
  TRY(dev_dup(s, d.lane_index, L)); TRY(dev_dup(s, d.lane_perm, L)); TRY(dev_dup(s, d.lane_internal, L));
  TRY(dev_dup(s, d.lane_left, L)); TRY(dev_dup(s, d.lane_right, L)); TRY(dev_dup(s, d.lane_link_off, L + 1));
  TRY(dev_dup(s, d.lane_tls_dist, L)); TRY(dev_dup(s, d.lane_sig, L)); TRY(dev_dup(s, d.lane_sig_slot, L));
  TRY(dev_dup(s, d.edge_lane0, sc->n_edges)); TRY(dev_dup(s, d.edge_nlanes, sc->n_edges));
  TRY(dev_dup(s, d.link_from, K)); TRY(dev_dup(s, d.link_to, K)); TRY(dev_dup(s, d.link_via, K));
  TRY(dev_dup(s, d.link_tls, K)); TRY(dev_dup(s, d.link_tlidx, K)); TRY(dev_dup(s, d.link_state, K));
  TRY(dev_dup(s, d.link_to_edge, K)); TRY(dev_dup(s, d.link_via_len, K)); TRY(dev_dup(s, d.link_last_int, K));
  TRY(dev_dup(s, d.link_cont, K)); TRY(dev_dup(s, d.link_parent, K)); TRY(dev_dup(s, d.link_foe_off, K + 1));
  TRY(dev_dup(s, d.foe_link, sc->n_foes)); TRY(dev_dup(s, d.foe_flags, sc->n_foes));
  TRY(dev_dup(s, d.tls_phase_off, sc->n_tls + 1)); TRY(dev_dup(s, d.tls_nlinks, sc->n_tls));
  TRY(dev_dup(s, d.tls_init_phase, sc->n_tls)); TRY(dev_dup(s, d.tls_init_left, sc->n_tls));
  TRY(dev_dup(s, d.phase_dur, sc->n_phases)); TRY(dev_dup(s, d.phase_state_off, sc->n_phases));
  TRY(dev_dup(s, d.state_chars, sc->n_state_chars));
  TRY(dev_dup(s, d.sig_tls, S)); TRY(dev_dup(s, d.sig_n_green, S)); TRY(dev_dup(s, d.sig_yellow_off, S + 1));
  TRY(dev_dup(s, d.yellow_idx, sc->n_yellow)); TRY(dev_dup(s, d.sig_lane_off, S + 1));
  TRY(dev_dup(s, d.sig_lane, sc->n_sig_lanes)); TRY(dev_dup(s, d.mv_off, S * 12 + 1));
  TRY(dev_dup(s, d.mv_lane, sc->n_mv_lanes)); TRY(dev_dup(s, d.mvo_off, S * 12 + 1));
  TRY(dev_dup(s, d.mvo_sig, sc->n_mvo)); TRY(dev_dup(s, d.mvo_slot, sc->n_mvo));
  TRY(dev_dup(s, d.out_off, S + 1)); TRY(dev_dup(s, d.out_sig, sc->n_out)); TRY(dev_dup(s, d.out_slot, sc->n_out));
  TRY(dev_dup(s, d.vtype, (size_t)sc->n_vtypes * 8)); TRY(dev_dup(s, d.vtype_bit, sc->n_vtypes));
  TRY(dev_dup(s, d.route_off, sc->n_routes + 1)); TRY(dev_dup(s, d.route_edge, sc->n_route_steps));
  TRY(dev_dup(s, d.route_mask, sc->n_route_steps));
  TRY(dev_dup(s, d.origin_lane, sc->n_origins)); TRY(dev_dup(s, d.origin_off, sc->n_origins + 1));
  TRY(dev_dup(s, d.trip_depart, sc->n_trips)); TRY(dev_dup(s, d.trip_route, sc->n_trips));
  TRY(dev_dup(s, d.trip_vtype, sc->n_trips)); TRY(dev_dup(s, d.trip_file, sc->n_trips));
  TRY(dev_dup(s, d.trip_depart_pos, sc->n_trips));
  TRY(dev_dup(s, d.origin_rate, sc->n_origins)); TRY(dev_dup(s, d.origin_route_off, sc->n_origins + 1));
  TRY(dev_dup(s, d.origin_route, sc->n_origin_routes));
  TRY(dev_dup(s, d.origin_watch_off, sc->n_origins + 1)); TRY(dev_dup(s, d.origin_watch_lane, sc->n_watch));
  TRY(dev_dup(s, d.origin_watch_dist, sc->n_watch)); TRY(dev_dup(s, d.origin_watch_owner, sc->n_watch));
  TRY(dev_dup(s, d.lane_watch_off, sc->n_lanes + 1)); TRY(dev_dup(s, d.lane_watch_lane, sc->n_lane_watch));
  TRY(dev_dup(s, d.lane_watch_dist, sc->n_lane_watch));
  const size_t N = (size_t)n_env;
  TRY(dev_alloc(s, s->d.hdr, N * kHdrInts));
  TRY(dev_alloc(s, s->d.tls_phase, N * sc->n_tls)); TRY(dev_alloc(s, s->d.tls_end, N * sc->n_tls));
  TRY(dev_alloc(s, s->d.next_phase, N * S));
  TRY(dev_alloc(s, s->d.origin_cur, N * sc->n_origins)); TRY(dev_alloc(s, s->d.origin_backlog, N * sc->n_origins));
  TRY(dev_alloc(s, s->d.veh, N * kVehWords * sc->vcap));
  const size_t SL = sc->n_sig_lanes;
  TRY(dev_alloc(s, s->d.lane_queue, N * SL)); TRY(dev_alloc(s, s->d.lane_approach, N * SL));
  TRY(dev_alloc(s, s->d.lane_total_wait, N * SL)); TRY(dev_alloc(s, s->d.lane_max_wait, N * SL));
  TRY(dev_alloc(s, s->d.lane_speed_sum, N * SL)); TRY(dev_alloc(s, s->d.lane_arrivals, N * SL));
  TRY(dev_alloc(s, s->d.phase_obs, N * S)); TRY(dev_alloc(s, s->d.mplight, N * S * 13));
  TRY(dev_alloc(s, s->d.wave, N * S * 12)); TRY(dev_alloc(s, s->d.rew_wait, N * S));
  TRY(dev_alloc(s, s->d.rew_wait_norm, N * S)); TRY(dev_alloc(s, s->d.rew_pressure, N * S));
  TRY(dev_alloc(s, s->d.sig_queue_len, N * S)); TRY(dev_alloc(s, s->d.sig_max_queue, N * S));
  s->d.drq = nullptr; s->d.drq_norm = nullptr; s->d.mplight_full = nullptr; s->d.out_mask = 0;   // allocated by rs_select_outputs
  s->d.trip_wait = nullptr;
  TRY(dev_alloc(s, s->d_actions, N * (S ? S : 1)));
  TRY(dev_alloc(s, s->d_stats, N));
  s->d.trip_rec = nullptr;
  if (sc->record_trips && !sc->synthetic && sc->n_trips > 0) {
    TRY(dev_alloc(s, s->d.trip_rec, N * (size_t)sc->n_trips)); TRY(dev_alloc(s, s->d.trip_wait, N * (size_t)sc->n_trips));
  }
  CK(cudaMallocHost((void**)&s->h_act_pinned, sizeof(int32_t) * N * (S ? S : 1)));
  s->host_obs_kind = RS_HOSTOBS_MPLIGHT; s->host_obs_floats = (size_t)(S ? S : 1) * 13;
  CK(cudaMallocHost((void**)&s->h_obs_pinned, sizeof(float) * N * s->host_obs_floats));
  CK(cudaMallocHost((void**)&s->h_rew_pinned, sizeof(float) * N * (S ? S : 1)));
  const char* eb = getenv("RESCO_B200_BLOCK");
  const char* er = getenv("RESCO_B200_REGCAP");
  const char* eg = getenv("RESCO_B200_GROUP");
  const char* es = getenv("RESCO_B200_SINGLE");   // 1 / 0 force the single-buffer tile on / off; default: automatic
  const char* egm = getenv("RESCO_B200_GMEM");    // 1: one pass with the whole store in the global-memory workspace
  const char* etl = getenv("RESCO_B200_TILE");    // vehicles in the fast pass's tile (overrides RsScenario.tile_vcap)
  const size_t optin = (size_t)prop.sharedMemPerBlockOptin;
  // largest compiled instances-per-CTA shape whose tiles fit the opt-in shared memory of one CTA
  auto fit_group = [&](size_t tile, int want) {
    int g = want;
    while (g > 1 && tile * g > optin) g = g > 8 ? 8 : (g > 4 ? g - 1 : g / 2);
    return g;
  };
  // ---- the vehicle store and the tile ----
  // `vcap` is the capacity of an instance's store in HBM (what SUMO does not have: keep it generous).  The env-step
  // kernel works on a TILE of the store: the fast pass holds `tile` vehicles per instance in shared memory, sized for
  // the traffic the map carries when it flows (about 1/24 of its jam capacity; cologne8: 128 vehicles, eight
  // instances per CTA).  An instance that outgrows the tile -- a jammed network -- is not truncated: the fast pass
  // leaves it untouched and the OVERFLOW PASS, the same kernel with the tile in a global-memory workspace (L2
  // resident) and room for the whole store, redoes its step.  Results do not depend on the tile size.
  const int store = sc->vcap;
  double jam = 0;
  for (int l = 0; l < L; ++l) if (!sc->lane_internal[l]) jam += floor(sc->lane_len[l] / 7.5) + 1.0;
  const bool force_gmem = egm && atoi(egm) != 0;
  int tile = etl ? atoi(etl) : sc->tile_vcap;
  if (tile <= 0) { tile = ((int)(jam / 24.0) + 31) / 32 * 32; tile = tile < 64 ? 64 : (tile > 1024 ? 1024 : tile); }
  tile = (tile + 3) / 4 * 4;
  if (tile > store || force_gmem) tile = store;
  s->d.sc.tile_cap = tile; s->d.sc.tile_single = 0; s->d.sc.tile_gmem = 0;
  s->d.from_list = 0; s->d.overflow_count = nullptr; s->d.overflow_list = nullptr; s->d.workspace = nullptr;
  s->layout = make_layout(s->d.sc);
  s->group = fit_group(s->layout.total, eg ? atoi(eg) : 8);
  s->minb = 1;
  // Big tiles (one or two instances per SM with the ping-pong tile): keep ONE tile buffer and run the per-tick
  // re-sort through registers (needs tile <= 2 x 512 threads).  Half the shared memory per instance doubles the
  // instances resident per SM, which is what hides the barrier waits of the plan phase on the big maps.
  const bool single_ok = tile <= 1024;
  const bool single = force_gmem ? false : ((eb && atoi(eb) < 512) ? false : (es ? (atoi(es) != 0 && single_ok) : (single_ok && s->group <= 2)));
  if (single) {
    s->d.sc.tile_single = 1;
    s->layout = make_layout(s->d.sc);
    // one instance per CTA and two independent CTAs per SM if both fit (1 KB per CTA is reserved by the driver; 64
    // registers per thread): measured 736 k env steps/s on ingolstadt21 (2048 instances) against 645 k for two
    // instances in lock-step in one CTA and 561 k for the ping-pong tile; grid4x4 synthetic 593 k / 508 k / 463 k
    const bool two_ctas = 2 * (s->layout.total + 1024) <= (size_t)prop.sharedMemPerMultiprocessor;
    s->group = fit_group(s->layout.total, eg ? atoi(eg) : (two_ctas ? 1 : 2));
    if (s->group == 1 && two_ctas) s->minb = 2;
  }
  // a tile that does not fit one CTA's shared memory at all (or RESCO_B200_GMEM=1): single pass out of the workspace
  const bool gmem = force_gmem || s->layout.total > optin;
  if (gmem) {
    s->d.sc.tile_single = 0; s->d.sc.tile_gmem = 1;
    s->layout = make_layout(s->d.sc);
    const bool two_ctas = 2 * (s->layout.total + 1024) <= (size_t)prop.sharedMemPerMultiprocessor;
    s->group = eg ? fit_group(s->layout.total, atoi(eg)) : 1;
    s->minb = (s->group == 1 && two_ctas) ? 2 : 1;
  }
  if (s->layout.total > optin) {
    char buf[256];
    snprintf(buf, sizeof buf, "rs_create: the per-lane tables of one instance need %zu B shared memory (> %zu B per CTA)",
             s->layout.total, optin);
    rs_destroy(s);
    return fail(RS_ERR_CAPACITY, buf);
  }
  // at least 512 threads per CTA whatever the tile size: a big map whose tile only fits once or twice per SM gets
  // 512 threads per instance, four times 128, instead of leaving the SM with two warps (measured on a B200,
  // ingolstadt21 2048 instances vcap 1024: 64 -> 126 k, 256 -> 283 k, 512 -> 404 k, 1024 -> 375 k env steps/s;
  // grid4x4 2 x 256 -> 363 k, 2 x 512 -> 377 k; cologne8 8 x 64 -> 3.55 M, 8 x 128 -> 2.53 M)
  s->block = eb ? atoi(eb) : (gmem ? 512 : (s->group >= 5 ? 64 : (s->group == 4 ? 128 : 512)));
  if (er) s->minb = atoi(er) != 0 ? 1024 / (s->block * s->group) : 1;
  if (getenv("RESCO_B200_MINB")) s->minb = atoi(getenv("RESCO_B200_MINB"));
  if (s->d.sc.tile_single && (s->block < 512 || tile > 2 * s->block)) {
    rs_destroy(s);
    return fail(RS_ERR_INVALID, "rs_create: the single-buffer tile needs tile_vcap <= 2 x threads per instance");
  }
  const char* ep = getenv("RESCO_B200_PERSIST");
  s->d.persistent = ep ? atoi(ep) : 1;
  const char* et = getenv("RESCO_B200_TMA");
  s->d.use_tma = et ? atoi(et) : 1;
  s->n_sm = prop.multiProcessorCount;
  TRY(dev_alloc(s, s->counters, 8));
  s->d.work_counter = s->counters;
  s->d.redo_count = s->counters + 3;
  TRY(dev_alloc(s, s->d.phase_clocks, 24));
  const char* ec = getenv("RESCO_B200_CARVEOUT");
  s->carveout = ec ? atoi(ec) : -1;
  const char* ex = getenv("RESCO_B200_SMEM_EXTRA");
  s->smem_extra = ex ? atoi(ex) : 0;
  // ---- instances that outgrow the tile ----
  // (1) launches with several instances per CTA: the CTA that hits such an instance steps it again at once with all its
  //     threads on a tile laid over the CTA's whole shared memory (no second launch, no tail: the other CTAs keep
  //     pulling groups); (2) what outgrows that too -- or any tile of a one-instance-per-CTA launch -- goes on the
  //     overflow list and is stepped by the overflow pass out of the global-memory workspace.
  s->d.sc.redo_cap = 0; s->d.sc.redo_single = 0;
  if (!gmem && s->group > 1 && tile < store && !(getenv("RESCO_B200_REDO") && atoi(getenv("RESCO_B200_REDO")) == 0)) {
    const size_t cta_smem = s->layout.total * s->group;
    const int threads = s->block * s->group;
    const int single1 = threads >= 512 ? 1 : 0;
    int cap = store;
    if (single1) { cap = cap < 1024 ? cap : 1024; cap = cap < 2 * threads ? cap : 2 * threads; }
    cap = cap / 4 * 4;
    while (cap > tile && make_layout_ex(s->d.sc, cap, single1, 0).total > cta_smem) cap -= 32;
    if (cap > tile) { s->d.sc.redo_cap = cap; s->d.sc.redo_single = single1; }
  }
  s->two_pass = !gmem && (s->d.sc.redo_cap > 0 ? s->d.sc.redo_cap : tile) < store;
  s->d.heavy_count = nullptr; s->d.heavy_list[0] = s->d.heavy_list[1] = nullptr; s->d.heavy_taken = s->counters + 4;
  s->d.heavy_thr = tile - 16;
  if (s->d.sc.redo_cap > 0 && !(getenv("RESCO_B200_HEAVY") && atoi(getenv("RESCO_B200_HEAVY")) == 0)) {
    s->d.persistent = 1;
    TRY(dev_alloc(s, s->d.heavy_list[0], N)); TRY(dev_alloc(s, s->d.heavy_list[1], N));
    TRY(dev_alloc(s, s->d.heavy_count, 4));
  }
  s->block2 = 512; s->group2 = 1; s->minb2 = 1;
  if (s->two_pass) {
    s->d.persistent = 1;
    DevScenario sc2 = s->d.sc;
    sc2.tile_cap = store; sc2.tile_single = 0; sc2.tile_gmem = 1;
    s->layout2 = make_layout(sc2);
    if (s->layout2.total > optin) { rs_destroy(s); return fail(RS_ERR_CAPACITY, "rs_create: per-lane tables exceed the shared memory of one CTA"); }
    s->minb2 = 2 * (s->layout2.total + 1024) <= (size_t)prop.sharedMemPerMultiprocessor ? 2 : 1;
    TRY(dev_alloc(s, s->d.overflow_list, N));
    s->d.overflow_count = s->counters + 2;
  }
  TRY(configure(s));
  {
    int grid = (n_env + s->group - 1) / s->group;
    if (s->resident_ctas < grid) grid = s->resident_ctas;
    if (gmem) {
      s->d.persistent = 1;   // the workspace is sized for the resident grid
      TRY(dev_alloc(s, s->d.workspace, (size_t)grid * s->group * s->layout.veh_total));
    }
    if (s->two_pass) {
      int grid2 = n_env < s->resident_ctas2 ? n_env : s->resident_ctas2;
      TRY(dev_alloc(s, s->workspace2, (size_t)grid2 * s->layout2.veh_total));
    }
  }
  CK(cudaEventCreate(&s->ev0)); CK(cudaEventCreate(&s->ev1));
  CK(cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming));
  *out = s;
  return rs_reset(s, seed, 0, nullptr);
}

extern "C" int rs_destroy(RsSim* s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  for (void* p : s->allocs) cudaFree(p);
  if (s->h_act_pinned) cudaFreeHost(s->h_act_pinned);
  if (s->h_obs_pinned) cudaFreeHost(s->h_obs_pinned);
  if (s->h_rew_pinned) cudaFreeHost(s->h_rew_pinned);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->ev_done) cudaEventDestroy(s->ev_done);
  if (s->graph) cudaGraphExecDestroy(s->graph);
  delete s;
  return 0;
}

extern "C" int rs_reset(RsSim* s, uint64_t seed, int64_t first_env_id, void* stream) {
  if (!s) return fail(RS_ERR_INVALID, "rs_reset: null sim");
  CK(cudaSetDevice(s->device));
  s->d.seed = seed; s->d.first_env_id = first_env_id;
  s->graph_policy = 0;   // the captured kernels carry the old seed
  k_reset<<<s->d.n_env, 64, 0, (cudaStream_t)stream>>>(s->d);
  if (s->d.heavy_count) CK(cudaMemsetAsync(s->d.heavy_count, 0, 4 * sizeof(int32_t), (cudaStream_t)stream));   // empty network: nobody is heavy
  s->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int rs_set_demand_window(RsSim* s, const int32_t* h_origin_off) {
  if (!s || !h_origin_off) return fail(RS_ERR_INVALID, "rs_set_demand_window: bad arguments");
  const int O = s->d.sc.n_origins;
  if (s->d.sc.synthetic) return fail(RS_ERR_INVALID, "rs_set_demand_window: the scenario has synthetic demand, not a trip table");
  for (int o = 0; o < O; ++o)
    if (h_origin_off[o] < 0 || h_origin_off[o] > h_origin_off[o + 1] || h_origin_off[o + 1] > s->d.sc.n_trips)
      return fail(RS_ERR_INVALID, "rs_set_demand_window: range outside the trip table");
  CK(cudaSetDevice(s->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(const_cast<int32_t*>(s->d.sc.origin_off), h_origin_off, sizeof(int32_t) * (O + 1), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int rs_set_phase(RsSim* s, const int32_t* d_phase, const uint8_t* d_mask, void* stream) {
  if (!s || !d_phase) return fail(RS_ERR_INVALID, "rs_set_phase: bad arguments");
  int total = s->d.n_env * s->d.sc.n_signals;
  if (total == 0) return 0;
  k_set_phase<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(s->d, d_phase, d_mask);
  s->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int rs_tick(RsSim* s, int32_t n_ticks, void* stream) {
  if (!s || n_ticks < 0) return fail(RS_ERR_INVALID, "rs_tick: bad arguments");
  RunArgs a{nullptr, 0, n_ticks, 0, 0, 0};
  return run(s, a, (cudaStream_t)stream);
}

extern "C" int rs_observe(RsSim* s, void* stream) {
  if (!s) return fail(RS_ERR_INVALID, "rs_observe: null sim");
  RunArgs a{nullptr, 0, 0, 0, 0, 1};
  return run(s, a, (cudaStream_t)stream);
}

extern "C" int rs_env_step(RsSim* s, const int32_t* d_actions, void* stream) {
  if (!s || !d_actions) return fail(RS_ERR_INVALID, "rs_env_step: bad arguments");
  const DevScenario& sc = s->d.sc;
  RunArgs a{d_actions, 1, sc.yellow_length, 1, sc.step_length - sc.yellow_length, 1};
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaEventRecord(s->ev0, st));
  int r = run(s, a, st);
  if (r) return r;
  CK(cudaEventRecord(s->ev1, st));
  s->timed = true;
  return 0;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

extern "C" int rs_env_step_host_async(RsSim* s, const int32_t* h_actions, float* h_obs, float* h_reward,
                                      int32_t reward_kind, void* stream) {
  if (!s || !h_actions) return fail(RS_ERR_INVALID, "rs_env_step_host_async: bad arguments");
  if (s->pending) return fail(RS_ERR_INVALID, "rs_env_step_host_async: a step is already pending (call rs_wait)");
  CK(cudaSetDevice(s->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t NS = (size_t)s->d.n_env * s->d.sc.n_signals;
  // page-locked caller buffers are used directly; pageable ones go through the sim's pinned staging buffers
  const bool pa = is_pinned(h_actions), po = !h_obs || is_pinned(h_obs), pr = !h_reward || is_pinned(h_reward);
  const int32_t* src = h_actions;
  if (!pa) { memcpy(s->h_act_pinned, h_actions, sizeof(int32_t) * NS); src = s->h_act_pinned; }
  CK(cudaMemcpyAsync(s->d_actions, src, sizeof(int32_t) * NS, cudaMemcpyHostToDevice, st));
  int r = rs_env_step(s, s->d_actions, st);
  if (r) return r;
  const float* rew = reward_kind == 0 ? s->d.rew_wait : (reward_kind == 1 ? s->d.rew_wait_norm : s->d.rew_pressure);
  const size_t obs_bytes = sizeof(float) * (size_t)s->d.n_env * s->host_obs_floats;
  const float* obs_src = s->host_obs_kind == RS_HOSTOBS_WAVE ? s->d.wave : s->host_obs_kind == RS_HOSTOBS_DRQ_NORM ? s->d.drq_norm
                       : s->host_obs_kind == RS_HOSTOBS_DRQ ? s->d.drq : s->host_obs_kind == RS_HOSTOBS_MPLIGHT_FULL ? s->d.mplight_full
                       : s->d.mplight;
  if (h_obs) CK(cudaMemcpyAsync(po ? h_obs : s->h_obs_pinned, obs_src, obs_bytes, cudaMemcpyDeviceToHost, st));
  if (h_reward) CK(cudaMemcpyAsync(pr ? h_reward : s->h_rew_pinned, rew, sizeof(float) * NS, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(s->ev_done, st));
  s->pending = true;
  s->pend_obs = (h_obs && !po) ? h_obs : nullptr;
  s->pend_rew = (h_reward && !pr) ? h_reward : nullptr;
  return 0;
}

extern "C" int rs_wait(RsSim* s) {
  if (!s) return fail(RS_ERR_INVALID, "rs_wait: null sim");
  if (!s->pending) return 0;
  s->pending = false;
  CK(cudaEventSynchronize(s->ev_done));
  const size_t NS = (size_t)s->d.n_env * s->d.sc.n_signals;
  if (s->pend_obs) memcpy(s->pend_obs, s->h_obs_pinned, sizeof(float) * (size_t)s->d.n_env * s->host_obs_floats);
  if (s->pend_rew) memcpy(s->pend_rew, s->h_rew_pinned, sizeof(float) * NS);
  return 0;
}

extern "C" int rs_env_step_host(RsSim* s, const int32_t* h_actions, float* h_obs, float* h_reward, int32_t reward_kind) {
  int r = rs_env_step_host_async(s, h_actions, h_obs, h_reward, reward_kind, nullptr);
  return r ? r : rs_wait(s);
}

extern "C" int rs_policy_maxpressure(RsSim* s, const int32_t* h_pairs, int32_t n_pairs, const int32_t* h_valid,
                                     int32_t use_wave, int32_t* d_actions_out, void* stream) {
  if (!s || !h_pairs || !h_valid || n_pairs <= 0) return fail(RS_ERR_INVALID, "rs_policy_maxpressure: bad arguments");
  const int S = s->d.sc.n_signals;
  if (s->n_pairs_alloc != n_pairs) {
    TRY(dev_alloc(s, s->d_pairs, (size_t)n_pairs * 2));
    TRY(dev_alloc(s, s->d_valid, (size_t)n_pairs * S * 2));
    s->n_pairs_alloc = n_pairs;
    CK(cudaMemcpy(s->d_pairs, h_pairs, sizeof(int32_t) * n_pairs * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->d_valid, h_valid, sizeof(int32_t) * n_pairs * S * 2, cudaMemcpyHostToDevice));
  }
  int total = s->d.n_env * S;
  int32_t* outp = d_actions_out ? d_actions_out : s->d_actions;
  k_policy<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(s->d, s->d_pairs, n_pairs, s->d_valid, use_wave, outp);
  s->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int rs_frap_load(RsSim* s, const RsFrapParams* p, const int32_t* h_pairs, int32_t n_pairs, const int32_t* h_order) {
  if (!s || !p || !h_pairs || !h_order || n_pairs < 2 || n_pairs > kFrapMaxPairs) return fail(RS_ERR_INVALID, "rs_frap_load: bad arguments");
  CK(cudaSetDevice(s->device));
  const int S = s->d.sc.n_signals;
  std::vector<float> w(FP_TOTAL);
  struct { int off; const float* src; int n; } parts[] = {
      {FP_P, p->p_weight, 8}, {FP_DW, p->d_weight, 4}, {FP_DB, p->d_bias, 4}, {FP_LEW, p->lane_embedding_weight, 128},
      {FP_LEB, p->lane_embedding_bias, 16}, {FP_LCW, p->lane_conv_weight, 640}, {FP_LCB, p->lane_conv_bias, 20},
      {FP_REW, p->relation_embedding_weight, 8}, {FP_RCW, p->relation_conv_weight, 80}, {FP_RCB, p->relation_conv_bias, 20},
      {FP_HLW, p->hidden_layer_weight, 400}, {FP_HLB, p->hidden_layer_bias, 20}, {FP_BMW, p->before_merge_weight, 20},
      {FP_BMB, p->before_merge_bias, 1}};
  for (auto& q : parts) { if (!q.src) return fail(RS_ERR_INVALID, "rs_frap_load: null parameter tensor"); memcpy(w.data() + q.off, q.src, sizeof(float) * q.n); }
  // competition mask (agents/mplight.py:19-31): pair i and the j-th OTHER pair share exactly one movement
  std::vector<uint8_t> comp((size_t)n_pairs * (n_pairs - 1), 0);
  for (int i = 0; i < n_pairs; ++i) {
    int cnt = 0;
    for (int j = 0; j < n_pairs; ++j) {
      if (i == j) continue;
      int v[4] = {h_pairs[2 * i], h_pairs[2 * i + 1], h_pairs[2 * j], h_pairs[2 * j + 1]}, uniq = 0;
      for (int a = 0; a < 4; ++a) { bool dup = false; for (int b = 0; b < a; ++b) dup |= v[b] == v[a]; uniq += !dup; }
      comp[(size_t)i * (n_pairs - 1) + cnt++] = uniq == 3;
    }
  }
  for (int q = 0; q < 2 * n_pairs; ++q) if (h_pairs[q] < 0 || h_pairs[q] >= RS_N_MOVEMENTS) return fail(RS_ERR_INVALID, "rs_frap_load: movement index out of range");
  if (!s->d_frap || s->frap_pairs != n_pairs) {
    TRY(dev_alloc(s, s->d_frap, FP_TOTAL)); TRY(dev_alloc(s, s->d_frap_pairs, (size_t)n_pairs * 2));
    TRY(dev_alloc(s, s->d_frap_comp, comp.size())); TRY(dev_alloc(s, s->d_frap_order, (size_t)S * n_pairs * 2));
    s->frap_pairs = n_pairs;
    CK(cudaFuncSetAttribute(k_policy_frap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)frap_smem_bytes(n_pairs)));
  }
  CK(cudaMemcpy(s->d_frap, w.data(), sizeof(float) * FP_TOTAL, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s->d_frap_pairs, h_pairs, sizeof(int32_t) * n_pairs * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s->d_frap_comp, comp.data(), comp.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s->d_frap_order, h_order, sizeof(int32_t) * S * n_pairs * 2, cudaMemcpyHostToDevice));
  s->graph_policy = 0;
  return 0;
}

static int launch_frap(RsSim* s, const float* d_obs, int n_env_rows, int32_t* d_actions_out, float* d_q_out, cudaStream_t st) {
  const int S = s->d.sc.n_signals, n = s->frap_pairs;
  const int rows = (d_obs ? n_env_rows : s->d.n_env) * S;
  const int R = kFrapThreads / n;
  k_policy_frap<<<(rows + R - 1) / R, kFrapThreads, frap_smem_bytes(n), st>>>(d_obs ? d_obs : s->d.mplight, rows, S, s->d_frap,
      s->d_frap_pairs, n, s->d_frap_comp, s->d_frap_order, d_q_out, d_actions_out ? d_actions_out : s->d_actions);
  s->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int rs_policy_frap(RsSim* s, const float* d_obs, int32_t n_env_rows, int32_t* d_actions_out, float* d_q_out, void* stream) {
  if (!s || !s->d_frap || (d_obs && n_env_rows <= 0)) return fail(RS_ERR_INVALID, "rs_policy_frap: bad arguments (rs_frap_load first)");
  return launch_frap(s, d_obs, n_env_rows, d_actions_out, d_q_out, (cudaStream_t)stream);
}

extern "C" int rs_policy_random(RsSim* s, uint64_t seed, int32_t* d_actions_out, void* stream) {
  if (!s) return fail(RS_ERR_INVALID, "rs_policy_random: null sim");
  const int total = s->d.n_env * s->d.sc.n_signals;
  if (total == 0) return 0;
  k_policy_random<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(s->d, seed, d_actions_out ? d_actions_out : s->d_actions);
  s->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// policy kernel + fused env step captured once into a CUDA graph and replayed: one launch per agent step
extern "C" int rs_env_step_policy(RsSim* s, int32_t policy, uint64_t policy_seed, void* stream) {
  if (!s || policy < RS_POLICY_MAXPRESSURE || policy > RS_POLICY_RANDOM) return fail(RS_ERR_INVALID, "rs_env_step_policy: bad arguments");
  if ((policy == RS_POLICY_MAXPRESSURE || policy == RS_POLICY_MAXWAVE) && !s->n_pairs_alloc)
    return fail(RS_ERR_INVALID, "rs_env_step_policy: upload the action tables with one rs_policy_maxpressure call first");
  if (policy == RS_POLICY_FRAP && !s->d_frap) return fail(RS_ERR_INVALID, "rs_env_step_policy: rs_frap_load first");
  CK(cudaSetDevice(s->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (s->graph_policy != policy || s->graph_seed != policy_seed) {
    if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
    cudaStream_t cap;
    CK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    CK(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    const DevScenario& sc = s->d.sc;
    const int total = s->d.n_env * sc.n_signals;
    int r = 0;
    if (policy == RS_POLICY_FRAP) r = launch_frap(s, nullptr, 0, nullptr, nullptr, cap);
    else if (policy == RS_POLICY_RANDOM) { k_policy_random<<<(total + 127) / 128, 128, 0, cap>>>(s->d, policy_seed, s->d_actions); s->launches += 1; }
    else { k_policy<<<(total + 127) / 128, 128, 0, cap>>>(s->d, s->d_pairs, s->n_pairs_alloc, s->d_valid, policy == RS_POLICY_MAXWAVE, s->d_actions); s->launches += 1; }
    RunArgs a{s->d_actions, 1, sc.yellow_length, 1, sc.step_length - sc.yellow_length, 1};
    if (!r) r = run(s, a, cap);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(cap, &g);
    cudaStreamDestroy(cap);
    if (r) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) return fail(RS_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&s->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(RS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    s->graph_policy = policy; s->graph_seed = policy_seed;
  } else s->launches += 2;
  CK(cudaEventRecord(s->ev0, st));
  CK(cudaGraphLaunch(s->graph, st));
  CK(cudaEventRecord(s->ev1, st));
  s->timed = true;
  return 0;
}

// Host-side agent front-end: WaveAgent.act / MaxAgent.act (agents/maxwave.py:18-38,
// agents/maxpressure.py:13-18) over a batch of observation rows that already sit in HOST memory (the buffer
// rs_env_step_host / rs_wait filled).  This is the caller's side of the loop, not the simulation path: it
// touches no device state and needs no RsSim.  Same tables and tie rule (first maximum) as rs_policy_maxpressure.
extern "C" int rs_host_agent_wave(const float* h_obs, int32_t n_env, int32_t n_signals, int32_t obs_dim, int32_t skip,
                                  const int32_t* pairs, int32_t n_pairs, const int32_t* order, int32_t* h_actions) {
  if (!h_obs || !pairs || !order || !h_actions || n_env < 0 || n_signals <= 0 || n_pairs <= 0 || obs_dim <= 0 || skip < 0)
    return fail(RS_ERR_INVALID, "rs_host_agent_wave: bad arguments");
  for (int q = 0; q < n_pairs * 2; ++q)
    if (pairs[q] < 0 || pairs[q] + skip >= obs_dim) return fail(RS_ERR_INVALID, "rs_host_agent_wave: pair index outside the observation row");
  for (int64_t e = 0; e < n_env; ++e) {
    for (int sg = 0; sg < n_signals; ++sg) {
      const float* ob = h_obs + ((size_t)e * n_signals + sg) * obs_dim + skip;
      const int32_t* ord = order + (size_t)sg * n_pairs * 2;
      float best = 0.0f; int act = 0; bool have = false;
      for (int q = 0; q < n_pairs && ord[2 * q] >= 0; ++q) {
        const int pi = ord[2 * q];
        if (pi >= n_pairs) return fail(RS_ERR_INVALID, "rs_host_agent_wave: pair index out of range");
        const float press = ob[pairs[2 * pi]] + ob[pairs[2 * pi + 1]];
        if (!have || press > best) { best = press; act = ord[2 * q + 1]; have = true; }
      }
      h_actions[(size_t)e * n_signals + sg] = act;
    }
  }
  return 0;
}

extern "C" int rs_get_obs(RsSim* s, RsObsView* o) {
  if (!s || !o) return fail(RS_ERR_INVALID, "rs_get_obs: bad arguments");
  o->n_env = s->d.n_env; o->n_signals = s->d.sc.n_signals; o->n_sig_lanes = s->d.sc.n_sig_lanes;
  o->lane_queue = s->d.lane_queue; o->lane_approach = s->d.lane_approach; o->lane_total_wait = s->d.lane_total_wait;
  o->lane_max_wait = s->d.lane_max_wait; o->lane_speed_sum = s->d.lane_speed_sum; o->phase = s->d.phase_obs;
  o->mplight = s->d.mplight; o->wave = s->d.wave; o->reward_wait = s->d.rew_wait;
  o->reward_wait_norm = s->d.rew_wait_norm; o->reward_pressure = s->d.rew_pressure;
  o->sig_queue_len = s->d.sig_queue_len; o->sig_max_queue = s->d.sig_max_queue; o->lane_arrivals = s->d.lane_arrivals;
  o->drq = s->d.drq; o->drq_norm = s->d.drq_norm; o->mplight_full = s->d.mplight_full;
  return 0;
}

extern "C" int rs_get_stats(RsSim* s, RsStats* h_out) {
  if (!s || !h_out) return fail(RS_ERR_INVALID, "rs_get_stats: bad arguments");
  CK(cudaSetDevice(s->device));
  CK(cudaDeviceSynchronize());
  k_stats<<<(s->d.n_env + 63) / 64, 64>>>(s->d, s->d_stats);
  s->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpy(h_out, s->d_stats, sizeof(RsStats) * s->d.n_env, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int rs_dump_vehicles(RsSim* s, int32_t env, int32_t* n_out, int32_t* lane, float* pos, float* speed,
                                float* accel, float* wait, float* rwait, float* tloss, int32_t* vid, int32_t* vtype,
                                int32_t* route, int32_t* cursor, float* sf, int32_t* depart, float* acc_wait) {
  if (!s || env < 0 || env >= s->d.n_env || !n_out) return fail(RS_ERR_INVALID, "rs_dump_vehicles: bad arguments");
  CK(cudaSetDevice(s->device));
  CK(cudaDeviceSynchronize());
  const int vcap = s->d.sc.vcap;
  int32_t hdr[kHdrInts];
  CK(cudaMemcpy(hdr, s->d.hdr + (size_t)env * kHdrInts, sizeof hdr, cudaMemcpyDeviceToHost));
  int n = hdr[H_NVEH];
  std::vector<uint32_t> buf((size_t)kVehWords * vcap);
  CK(cudaMemcpy(buf.data(), s->d.veh + (size_t)env * kVehWords * vcap, buf.size() * 4, cudaMemcpyDeviceToHost));
  const float* fpos = (const float*)&buf[0]; const float* fspeed = (const float*)&buf[(size_t)vcap];
  const float* fsf = (const float*)&buf[2 * (size_t)vcap]; const float* ftl = (const float*)&buf[3 * (size_t)vcap];
  const uint32_t *wvid = &buf[4 * (size_t)vcap], *wr = &buf[5 * (size_t)vcap], *rc = &buf[6 * (size_t)vcap];
  const uint32_t *mt = &buf[7 * (size_t)vcap], *ed = &buf[8 * (size_t)vcap], *dl = &buf[9 * (size_t)vcap], *aw = &buf[10 * (size_t)vcap];
  for (int i = 0; i < n; ++i) {
    if (lane) lane[i] = (int32_t)(dl[i] >> 16);
    if (pos) pos[i] = fpos[i];
    if (speed) speed[i] = fspeed[i];
    if (accel) accel[i] = 0.0f;   // not part of the device state (TraCI facade differentiates speeds)
    if (wait) wait[i] = (float)(wr[i] & 0xFFFFu);
    if (rwait) rwait[i] = (float)(wr[i] >> 16);
    if (tloss) tloss[i] = ftl[i];
    if (vid) vid[i] = (int32_t)wvid[i];
    if (vtype) vtype[i] = (int32_t)(mt[i] & 0xFFu);
    if (route) route[i] = (int32_t)(rc[i] & 0xFFFFu);
    if (cursor) cursor[i] = (int32_t)(rc[i] >> 16);
    if (sf) sf[i] = fsf[i];
    if (depart) depart[i] = (int32_t)(ed[i] >> 16);
    if (acc_wait) acc_wait[i] = (float)(aw[i] & 0xFFFFu);
  }
  *n_out = n;
  return 0;
}

extern "C" int rs_get_phases(RsSim* s, int32_t env, int32_t* h_tls_phase) {
  if (!s || env < 0 || env >= s->d.n_env || !h_tls_phase) return fail(RS_ERR_INVALID, "rs_get_phases: bad arguments");
  CK(cudaSetDevice(s->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h_tls_phase, s->d.tls_phase + (size_t)env * s->d.sc.n_tls, sizeof(int32_t) * s->d.sc.n_tls,
                cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int rs_get_trip_records(RsSim* s, int32_t env, int32_t* h_arrival_tick, int32_t* h_depart_tick,
                                   float* h_time_loss, int32_t* h_depart_delay, float* h_waiting_time) {
  if (!s || env < 0 || env >= s->d.n_env) return fail(RS_ERR_INVALID, "rs_get_trip_records: bad arguments");
  if (!s->d.trip_rec) return fail(RS_ERR_INVALID, "rs_get_trip_records: RsScenario.record_trips was not set");
  CK(cudaSetDevice(s->device));
  CK(cudaDeviceSynchronize());
  const size_t n = (size_t)s->d.sc.n_trips;
  std::vector<int4> buf(n);
  CK(cudaMemcpy(buf.data(), s->d.trip_rec + (size_t)env * n, sizeof(int4) * n, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) {
    if (h_arrival_tick) h_arrival_tick[i] = buf[i].x;
    if (h_depart_tick) h_depart_tick[i] = buf[i].y;
    if (h_time_loss) memcpy(&h_time_loss[i], &buf[i].z, 4);
    if (h_depart_delay) h_depart_delay[i] = buf[i].w;
  }
  if (h_waiting_time) CK(cudaMemcpy(h_waiting_time, s->d.trip_wait + (size_t)env * n, sizeof(float) * n, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int rs_select_outputs(RsSim* s, int32_t mask) {
  if (!s || (mask & ~(RS_OUT_DRQ | RS_OUT_DRQ_NORM | RS_OUT_MPLIGHT_FULL))) return fail(RS_ERR_INVALID, "rs_select_outputs: bad arguments");
  CK(cudaSetDevice(s->device));
  const size_t N = (size_t)s->d.n_env, SL = (size_t)s->d.sc.n_sig_lanes, S = (size_t)s->d.sc.n_signals;
  if ((mask & RS_OUT_DRQ) && !s->d.drq) TRY(dev_alloc(s, s->d.drq, N * SL * 5));
  if ((mask & RS_OUT_DRQ_NORM) && !s->d.drq_norm) TRY(dev_alloc(s, s->d.drq_norm, N * SL * 5));
  if ((mask & RS_OUT_MPLIGHT_FULL) && !s->d.mplight_full) TRY(dev_alloc(s, s->d.mplight_full, N * S * 49));
  s->d.out_mask = mask;
  s->graph_policy = 0;
  return 0;
}

extern "C" int rs_set_host_obs(RsSim* s, int32_t kind, int32_t* floats_per_instance) {
  if (!s || kind < RS_HOSTOBS_MPLIGHT || kind > RS_HOSTOBS_MPLIGHT_FULL) return fail(RS_ERR_INVALID, "rs_set_host_obs: bad arguments");
  if (s->pending) return fail(RS_ERR_INVALID, "rs_set_host_obs: a step is pending (call rs_wait)");
  const size_t SL = (size_t)s->d.sc.n_sig_lanes, S = (size_t)s->d.sc.n_signals;
  const size_t fl = kind == RS_HOSTOBS_MPLIGHT ? S * 13 : kind == RS_HOSTOBS_WAVE ? S * 12 : kind == RS_HOSTOBS_MPLIGHT_FULL ? S * 49 : SL * 5;
  const int need = kind == RS_HOSTOBS_DRQ_NORM ? RS_OUT_DRQ_NORM : kind == RS_HOSTOBS_DRQ ? RS_OUT_DRQ : kind == RS_HOSTOBS_MPLIGHT_FULL ? RS_OUT_MPLIGHT_FULL : 0;
  if (need) TRY(rs_select_outputs(s, s->d.out_mask | need));
  if (fl > s->host_obs_floats) {   // the staging buffer for pageable callers grows with the row size
    if (s->h_obs_pinned) cudaFreeHost(s->h_obs_pinned);
    s->h_obs_pinned = nullptr;
    CK(cudaMallocHost((void**)&s->h_obs_pinned, sizeof(float) * (size_t)s->d.n_env * (fl ? fl : 1)));
  }
  s->host_obs_kind = kind; s->host_obs_floats = fl ? fl : 1;
  if (floats_per_instance) *floats_per_instance = (int32_t)fl;
  return 0;
}

extern "C" int64_t rs_kernel_launches(RsSim* s) { return s ? s->launches : 0; }

// Diagnostics: cycles per kernel phase summed over CTAs since rs_create (all zero unless the library was built with
// -DRS_PHASE_CLOCKS=1; see tools/phase_clocks.sh).  Not part of include/resco_b200.h.
extern "C" int rs_debug_phase_clocks(RsSim* s, unsigned long long* h_out24) {
  if (!s || !h_out24) return fail(RS_ERR_INVALID, "rs_debug_phase_clocks: bad arguments");
  CK(cudaSetDevice(s->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h_out24, s->d.phase_clocks, 24 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return RS_PHASE_CLOCKS;
}

extern "C" int rs_get_launch_shape(RsSim* s, int32_t* threads_per_instance, int32_t* instances_per_cta, int32_t* grid_ctas,
                                   int32_t* smem_bytes_per_cta, int32_t* tile_buffers) {
  if (!s) return fail(RS_ERR_INVALID, "rs_get_launch_shape: null sim");
  int grid = (s->d.n_env + s->group - 1) / s->group;
  if (s->d.persistent && s->resident_ctas < grid) grid = s->resident_ctas;
  if (threads_per_instance) *threads_per_instance = s->block;
  if (instances_per_cta) *instances_per_cta = s->group;
  if (grid_ctas) *grid_ctas = grid;
  if (smem_bytes_per_cta) *smem_bytes_per_cta = (int32_t)(s->layout.total * s->group);
  if (tile_buffers) *tile_buffers = s->d.sc.tile_gmem ? 0 : (s->d.sc.tile_single ? 1 : 2);   // 0: tile in the global workspace
  return 0;
}

extern "C" int rs_get_tile_info(RsSim* s, int32_t* tile_vcap, int32_t* store_vcap, int32_t* redo_vcap, int32_t* has_overflow_pass,
                                int32_t* last_redone, int32_t* last_deferred) {
  if (!s) return fail(RS_ERR_INVALID, "rs_get_tile_info: null sim");
  if (tile_vcap) *tile_vcap = s->d.sc.tile_cap;
  if (store_vcap) *store_vcap = s->d.sc.vcap;
  if (redo_vcap) *redo_vcap = s->d.sc.redo_cap;
  if (has_overflow_pass) *has_overflow_pass = s->two_pass ? 1 : 0;
  if (last_redone || last_deferred) {
    int32_t c[4] = {0, 0, 0, 0};
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(c, s->counters, sizeof c, cudaMemcpyDeviceToHost));
    if (last_redone) *last_redone = c[3];
    if (last_deferred) *last_deferred = c[2];
  }
  return 0;
}

extern "C" int rs_last_step_ms(RsSim* s, float* ms) {
  if (!s || !ms || !s->timed) return fail(RS_ERR_INVALID, "rs_last_step_ms: no timed step");
  CK(cudaEventSynchronize(s->ev1));
  CK(cudaEventElapsedTime(ms, s->ev0, s->ev1));
  return 0;
}
