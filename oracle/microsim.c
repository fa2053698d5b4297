/*
 * oracle/microsim.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product path (resco_b200/csrc) never calls it and has no CPU fallback.
 *
 * PARITY STATUS: **parity unpinned for the simulator core** -- the arithmetic of the
 * per-second step lives in Eclipse SUMO (pip `traci`/`libsumo`, unpinned in the reference's
 * setup.py:24-35; authors tested 1.9.0/1.9.1, README.md:7), whose source is NOT under
 * /root/reference and which is not installed here.  The car-following / junction / insertion
 * rules below restate SUMO's published algorithm (Krauss 1998 as implemented by SUMO's
 * MSCFModel / MSCFModel_Krauss Euler update, "Definition of Vehicles, Vehicle Types and Routes"
 * and "Simulation/Intersections" documentation) in single precision, from memory of the public
 * documentation; they have never been diffed against a libsumo trace.
 * PINNED against the reference's own Python (run here through a stub TraCI, see
 * tools/make_golden.py): create_yellows (traffic_signal.py:7-24), the phase machine
 * prep_phase/set_phase (:176-187), Signal.observe incl. the waiting-time latch (:189-235),
 * get_vehicles detector range (:238-247), states.mplight/wave/drq_norm (states.py:34-80,116-127),
 * rewards.wait/wait_norm/pressure (rewards.py:6-41), calc_metrics (multi_signal.py:199-216) and
 * the env-step schedule (multi_signal.py:164-197).
 *
 * Structure: plain sequential loops over one instance at a time.  Every decision in a tick reads
 * the state at the START of the tick (Jacobi style), which is what makes the rule set
 * order-independent and lets the CUDA path reproduce it bit for bit.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, plain IEEE f32).
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/resco_b200.h"

#define MAX_HOPS 8
#define HALT_SPEED 0.1f
#define NUM_EPS 0.001f
#define EMERGENCY_DECEL 9.0f
#define LC_COOLDOWN 5
#define COOP_MARGIN 1.0f   /* a yielding follower leaves this much more than minGap behind an urgent lane changer, so that the
                              changer's own safety test (gap >= follower speed x 1 s) passes while the follower still creeps */

typedef struct {
  float pos, speed, accel, sf, wait, rwait, tloss, await;   /* await: tripinfo waitingTime (all seconds with v < 0.1) */
  int32_t vid, vtype, route, cursor, depart, ddelay, seen_epoch, seen_sig, lcc;
} Veh;

typedef struct {
  int32_t tick, n_veh, epoch;
  Veh* veh;            /* CSR order: lane-major, front (largest pos) first */
  Veh* veh2;           /* scratch for the rebuild */
  int32_t* lane_start; /* [L+1] */
  int32_t* lane_start2;
  int32_t* tls_phase;  /* [n_tls] */
  int32_t* tls_end;    /* tick at which the phase expires */
  int32_t* next_phase; /* [S] */
  int32_t* origin_cur; /* [O] cursor into trips (table) / serial (synthetic) */
  int32_t* origin_backlog;
  /* per-tick scratch */
  float* tail_back;    /* [L] */
  float* tail_speed;
  float* tail_decel;
  float* lane_occ;     /* sum of (length + minGap) of the vehicles on the lane */
  float* vnext;        /* [vcap] */
  int32_t* new_lane;   /* [vcap]: -1 arrived */
  float* new_pos;
  int32_t* new_cursor;
  /* observation */
  float *lane_queue, *lane_approach, *lane_total_wait, *lane_max_wait, *lane_speed_sum, *lane_arrivals;
  int32_t* phase_obs;
  float *mplight, *wave, *rew_wait, *rew_wait_norm, *rew_pressure;
  float *drq, *drq_norm, *mplight_full;
  int32_t *sig_queue_len, *sig_max_queue;
  RsStats st;
  uint64_t env_id;
  int32_t *trip_arrival, *trip_depart_tick, *trip_ddelay; float *trip_tloss, *trip_wait;   /* [n_trips] when record_trips */
} Inst;

typedef struct OrcSim {
  RsScenario sc; /* deep copy */
  int32_t n_env;
  uint64_t seed;
  Inst* inst;
  void** owned;
  int n_owned;
} OrcSim;

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al. 2011) */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    if (r > 0) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
}
void orc_philox(uint32_t* ctr_inout, uint32_t k0, uint32_t k1) { philox(ctr_inout, k0, k1); }

enum { STREAM_SF = 1, STREAM_DAWDLE = 2, STREAM_DEMAND = 3, STREAM_ROUTE = 4, STREAM_DEPARTPOS = 5 };

static void rng4(const OrcSim* s, const Inst* in, uint32_t stream, uint32_t a, uint32_t b, uint32_t out[4]) {
  out[0] = (uint32_t)in->env_id; out[1] = (uint32_t)(in->env_id >> 32); out[2] = a; out[3] = b;
  philox(out, (uint32_t)s->seed ^ (stream * 0x632BE5ABu), (uint32_t)(s->seed >> 32));
}

/* speedFactor ~ N(1, dev) via a sum of eight 16-bit uniforms (exact integer arithmetic; SUMO draws
 * normc(1, dev, 0.2, 2) -- same first two moments, same clipping). */
static float speed_factor(const OrcSim* s, const Inst* in, int32_t vid, float dev) {
  if (!(dev > 0.0f)) return 1.0f;
  uint32_t r[4];
  rng4(s, in, STREAM_SF, (uint32_t)vid, 0u, r);
  uint32_t sum = 0;
  for (int i = 0; i < 4; ++i) sum += (r[i] & 0xFFFFu) + (r[i] >> 16);
  float z = ((float)sum - 262140.0f) * (1.0f / 65536.0f) * 1.2247449f;
  float sf = 1.0f + dev * z;
  return fminf(fmaxf(sf, 0.2f), 2.0f);
}

/* ------------------------------------------------------------------ car-following primitives (Euler, dt = 1 s) */
static float brake_gap(float speed, float decel, float headway) {
  int steps = (int)(speed / decel);
  float fs = (float)steps;
  float a = fs * speed;
  float b = decel * fs;
  float c = b * (float)(steps + 1);
  float d = c / 2.0f;
  return (a - d) + speed * headway;
}

static float max_safe_stop_speed(float gap, float decel, float tau) {
  float g = gap - NUM_EPS;
  if (g < 0.0f) return 0.0f;
  float b = decel, t = tau;
  float q = 2.0f * g / b - t;
  float inner = 1.0f + 4.0f * (q + t * t);
  float n = floorf(0.5f - ((t + sqrtf(inner) * -0.5f)));
  float h = 0.5f * n * (n - 1.0f) * b + n * b * t;
  float r = (g - h) / (n + t);
  return n * b + r;
}

static float follow_speed(float gap, float vlead, float dlead, float decel, float tau) {
  float bg = brake_gap(vlead, fmaxf(decel, dlead), 0.0f);
  return max_safe_stop_speed(gap + bg, decel, tau);
}

static float free_speed(float decel, float dist, float target) {
  if (dist < target) return target;
  float b = decel;
  float bb = b + 2.0f * target;
  float y = fmaxf(0.0f, ((sqrtf(bb * bb + 8.0f * b * dist) - b) * 0.5f - target) / b);
  float yf = floorf(y);
  float exact = (yf * yf * b + yf * b) / 2.0f + yf * target + (y > yf ? target : 0.0f);
  return fmaxf(0.0f, dist - exact) / (yf + 1.0f) + yf * b + target;
}

static float dawdle(float v, float accel, float sigma, float xi) {
  if (v < accel) v -= sigma * v * xi; else v -= sigma * accel * xi;
  return fmaxf(0.0f, v);
}

/* ------------------------------------------------------------------ network helpers */
#define VT(s, i, f) ((s)->sc.vtype[(i) * 8 + (f)])
enum { VT_LEN = 0, VT_GAP, VT_ACCEL, VT_DECEL, VT_TAU, VT_SIGMA, VT_VMAX, VT_DEV };

static int choose_link(const RsScenario* sc, int lane, int route, int cursor) {
  int k0 = sc->lane_link_off[lane], k1 = sc->lane_link_off[lane + 1];
  if (sc->lane_internal[lane]) return k0 < k1 ? k0 : -2;
  int ro = sc->route_off[route], rn = sc->route_off[route + 1] - ro;
  if (cursor + 1 >= rn) return -1; /* route ends on this edge */
  int ne = sc->route_edge[ro + cursor + 1], mask = sc->route_mask[ro + cursor + 1];
  int best = -2, rank = 0;         /* -2: this lane does not lead on (wrong lane) */
  for (int k = k0; k < k1; ++k) {  /* prefer a target lane that is "best", then "ok", then any */
    if (sc->link_to_edge[k] != ne) continue;
    int ti = sc->lane_index[sc->link_to[k]];
    int r = ((mask >> (8 + ti)) & 1) ? 3 : (((mask >> ti) & 1) ? 2 : 1);
    if (r > rank) { rank = r; best = k; }
  }
  return best;
}

static int state_now(const OrcSim* s, const Inst* in, int k) {
  const RsScenario* sc = &s->sc;
  int t = sc->link_tls[k];
  if (t < 0) return sc->link_state[k];
  int p = sc->tls_phase_off[t] + in->tls_phase[t];
  return sc->state_chars[sc->phase_state_off[p] + sc->link_tlidx[k]];
}

static int lane_count(const Inst* in, int l) { return in->lane_start[l + 1] - in->lane_start[l]; }

static int time_conflict(float seen, float v, float cross, float dist_f, float v_f, float cross_f) {
  float vm = fmaxf(v, 4.0f), vf = fmaxf(v_f, 2.0f);
  float tm_a = seen / vm, tm_l = (seen + cross) / vm;
  float tf_a = dist_f / vf, tf_l = (dist_f + cross_f) / vf;
  return (tf_a < tm_l + 1.0f) && (tf_l + 1.0f > tm_a);
}

/* must the vehicle on entry link k (distance `seen` to its stop line) wait for a foe? */
static int link_blocked(const OrcSim* s, const Inst* in, int k, float seen, float v, float cross) {
  const RsScenario* sc = &s->sc;
  for (int fi = sc->link_foe_off[k]; fi < sc->link_foe_off[k + 1]; ++fi) {
    int f = sc->foe_link[fi], fl = sc->foe_flags[fi];
    int li = sc->link_last_int[f];
    if (li >= 0 && lane_count(in, li) > 0) return 100000 + f; /* somebody is crossing my path */
    if (!(fl & 1)) continue;                         /* I have right of way over f */
    int a0 = sc->link_from[f];
    if (lane_count(in, a0) > 0) {
      const Veh* h = &in->veh[in->lane_start[a0]];
      if (choose_link(sc, a0, h->route, h->cursor) == f) {
        float dist_f = sc->lane_len[a0] - h->pos;
        int fst = state_now(s, in, f), goes = 1;
        if (fst == 'r' || fst == 'u' || fst == 's') goes = 0;
        else if (fst == 'y' && dist_f >= brake_gap(h->speed, VT(s, h->vtype, VT_DECEL), 0.0f)) goes = 0;
        if (h->speed < HALT_SPEED) goes = 0; /* a standing foe is not approaching (it re-registers once it moves) */
        if (goes && time_conflict(seen, v, cross, dist_f, h->speed,
                                  sc->link_via_len[f] + VT(s, h->vtype, VT_LEN))) return 200000 + f;
      }
    }
    if (sc->link_cont[f]) {
      int a1 = sc->link_via[f];
      if (lane_count(in, a1) > 0) {
        const Veh* h = &in->veh[in->lane_start[a1]];
        float dist_f = sc->lane_len[a1] - h->pos;
        int goes = h->speed >= HALT_SPEED;
        if (goes && time_conflict(seen, v, cross, dist_f, h->speed,
                                  sc->link_via_len[f] - sc->lane_len[a1] + VT(s, h->vtype, VT_LEN))) return 300000 + f;
      }
    }
  }
  return 0;
}

/* stop-line decision for link k seen `seen` metres ahead by vehicle x (hop 0 = the link at the end of
 * its own lane) */
static int must_stop(const OrcSim* s, const Inst* in, const Veh* x, int k, float seen, int hop, int cursor, int binds) {
  const RsScenario* sc = &s->sc;
  float len = VT(s, x->vtype, VT_LEN), decel = VT(s, x->vtype, VT_DECEL);
  int from = sc->link_from[k];
  if (sc->lane_internal[from]) {
    int p = sc->link_parent[k];
    if ((hop == 0 || binds) && p >= 0 && sc->link_cont[p] && sc->link_via[p] == from)
      return link_blocked(s, in, p, seen, x->speed, sc->link_via_len[p] - sc->lane_len[from] + len);
    return 0;
  }
  int st = state_now(s, in, k);
  if (st == 'r' || st == 'u') return 1;
  if (st == 'y' || st == 'Y') return seen >= brake_gap(x->speed, decel, 0.0f) ? 2 : 0;
  if (st == 's' && !(x->wait > 0.0f && seen <= 2.0f)) return 3;
  /* right of way / keep-clear for a link further ahead only if stopping in front of it would bind the speed
   * now (short lanes are crossed within one tick, so the link at the end of the own lane is not enough) */
  if (hop != 0 && !binds) return 0;
  int minor = (st == 'g' || st == 'm' || st == '=' || st == 'Z' || st == 'w' || st == 's' || st == 'o');
  if (sc->link_cont[k]) {
    /* waiting slot inside the junction is taken by a STANDING vehicle (a moving one is simply followed) */
    int vl = sc->link_via[k];
    if (lane_count(in, vl) > 0 && in->veh[in->lane_start[vl + 1] - 1].speed < HALT_SPEED) return 4;
  } else if (minor) {
    int b = link_blocked(s, in, k, seen, x->speed, sc->link_via_len[k] + len);
    if (b) return b;
  } else {
    for (int fi = sc->link_foe_off[k]; fi < sc->link_foe_off[k + 1]; ++fi) {
      int li = sc->link_last_int[sc->foe_link[fi]];
      if (li >= 0 && lane_count(in, li) > 0) return 400000 + sc->foe_link[fi];
    }
  }
  /* keep the junction clear (SUMO MSVehicle::keepClear / checkRewindLinkLanes / MSLane::getSpaceTillLastStanding):
   * only links with foes, and only when a vehicle was seen beyond the stop line.  Space = room behind the last
   * STANDING vehicle of the lanes ahead (moving vehicles only take their own length), minus the vehicles that are
   * already inside this junction on my path; lanes are added up to the first standing vehicle / red light. */
  if (sc->link_foe_off[k + 1] > sc->link_foe_off[k]) {
    float need = len + VT(s, x->vtype, VT_GAP), space = 0.0f;
    int had = 0, cc2 = cursor + 1;
    int cur = sc->link_via[k] >= 0 ? sc->link_via[k] : sc->link_to[k];
    for (int h = 0; h < 3 && sc->lane_internal[cur]; ++h) {     /* my own path through the junction */
      if (lane_count(in, cur) > 0) { had = 1; space -= in->lane_occ[cur]; }
      int k2 = sc->lane_link_off[cur];
      cur = sc->link_via[k2] >= 0 ? sc->link_via[k2] : sc->link_to[k2];
    }
    for (int h = 0; h < 6; ++h) {
      int a = in->lane_start[cur], j = in->lane_start[cur + 1] - 1, stopped = 0;
      float lengths = 0.0f;
      if (j >= a) had = 1;
      for (; j >= a; --j) {                                      /* from the tail forward */
        const Veh* y = &in->veh[j];
        if (y->speed < HALT_SPEED) { stopped = 1; break; }
        lengths += VT(s, y->vtype, VT_LEN) + VT(s, y->vtype, VT_GAP);
      }
      if (stopped) { space += (in->veh[j].pos - VT(s, in->veh[j].vtype, VT_LEN)) - lengths; break; }
      if (!sc->lane_internal[cur]) space += sc->lane_len[cur] - lengths;   /* junction interiors are no place to stand */
      if (space >= need) return 0;
      int k2 = choose_link(sc, cur, x->route, cc2);
      if (k2 < 0) return 0;
      if (!sc->lane_internal[cur]) {
        int st2 = state_now(s, in, k2);
        if (st2 == 'r' || st2 == 'u' || st2 == 'y') break;
      }
      cur = sc->link_via[k2] >= 0 ? sc->link_via[k2] : sc->link_to[k2];
      if (!sc->lane_internal[cur]) cc2 += 1;
    }
    if (had && space < need) return 7;
  }
  return 0;
}

/* direction (+1 left / -1 right / 0) towards the nearest lane the route can continue from */
static int strategic_dir(const RsScenario* sc, const Veh* x, int lane) {
  int mask = sc->route_mask[sc->route_off[x->route] + x->cursor];
  int okm = mask & 0xFF, bestm = (mask >> 8) & 0xFF, myidx = sc->lane_index[lane];
  int want = !((okm >> myidx) & 1) ? okm : (!((bestm >> myidx) & 1) ? bestm : 0);
  if (!want) return 0;
  for (int d = 1; d < 8; ++d) {
    if (myidx + d < 8 && ((want >> (myidx + d)) & 1)) return 1;
    if (myidx - d >= 0 && ((want >> (myidx - d)) & 1)) return -1;
  }
  return 0;
}

/* ------------------------------------------------------------------ one tick */
typedef struct { float vsafe_lead; } PlanAux;

static void plan_vehicle(const OrcSim* s, Inst* in, int i, int lane, int rank) {
  const RsScenario* sc = &s->sc;
  const Veh* x = &in->veh[i];
  int vt = x->vtype;
  float len = VT(s, vt, VT_LEN), mingap = VT(s, vt, VT_GAP), accel = VT(s, vt, VT_ACCEL);
  float decel = VT(s, vt, VT_DECEL), tau = VT(s, vt, VT_TAU);
  float sigma = sc->sigma_override >= 0.0f ? sc->sigma_override : VT(s, vt, VT_SIGMA);
  float vcap = VT(s, vt, VT_VMAX);
  float v = x->speed;
  float vmaxl = fminf(sc->lane_vmax[lane] * x->sf, vcap);
  float vacc = fminf(v + accel, vmaxl);
  float vsafe = vacc;
  float vlead_limit = vacc; /* follow-speed w.r.t. the own-lane leader only (for the speed-gain rule) */
  int wrong_lane_head = 0;
  if (rank > 0) {
    const Veh* ld = &in->veh[i - 1];
    float gap = ld->pos - VT(s, ld->vtype, VT_LEN) - x->pos - mingap;
    vlead_limit = follow_speed(gap, ld->speed, VT(s, ld->vtype, VT_DECEL), decel, tau);
    vsafe = fminf(vsafe, vlead_limit);
  } else {
    float seen = sc->lane_len[lane] - x->pos;
    int cur = lane, cc = x->cursor;
    float la = brake_gap(vacc, decel, 0.0f) + 2.0f * vacc + 5.0f;
    /* a lane end further away than the look-ahead distance plus the longest vehicle that could still stick
     * out of the junction cannot bind the speed: no junction logic at all */
    const int far = seen > la + 20.0f;
    for (int hop = 0; hop < MAX_HOPS && !far; ++hop) {
      int k = choose_link(sc, cur, x->route, cc);
      if (k == -1) break;                       /* arrival at the end of this lane */
      if (k == -2) {
        vsafe = fminf(vsafe, max_safe_stop_speed(seen, decel, tau));
        if (hop == 0) wrong_lane_head = 1;
        break;
      }
      float vstop = max_safe_stop_speed(seen, decel, tau);
      if (must_stop(s, in, x, k, seen, hop, cc, vstop < vsafe)) { vsafe = fminf(vsafe, vstop); break; }
      int nxt = sc->link_via[k] >= 0 ? sc->link_via[k] : sc->link_to[k];
      vsafe = fminf(vsafe, free_speed(decel, seen, fminf(sc->lane_vmax[nxt] * x->sf, vcap)));
      if (lane_count(in, nxt) > 0) {
        float gap = seen + in->tail_back[nxt] - mingap;
        float f = follow_speed(gap, in->tail_speed[nxt], in->tail_decel[nxt], decel, tau);
        vsafe = fminf(vsafe, f);
        if (hop == 0) vlead_limit = fminf(vlead_limit, f);
        break;
      }
      seen += sc->lane_len[nxt];
      if (!sc->lane_internal[nxt]) cc += 1;
      cur = nxt;
      if (seen > la) break;
    }
  }
  /* ---- cooperation (LC2013 informFollower analogue): a vehicle of the neighbouring lane that MUST get into
   * this lane (its lane does not lead on) and is urgent becomes a virtual leader for everybody behind it ---- */
  if (sc->lane_change && !sc->lane_internal[lane]) {
    for (int side = 0; side < 2; ++side) {
      int nl = side == 0 ? sc->lane_left[lane] : sc->lane_right[lane];
      if (nl < 0) continue;
      int a = in->lane_start[nl], j = in->lane_start[nl + 1] - 1;
      float gapu = 0.0f;
      for (; j >= a; --j) {               /* from the tail forward: first vehicle that is entirely ahead of me */
        gapu = in->veh[j].pos - VT(s, in->veh[j].vtype, VT_LEN) - x->pos - mingap;
        if (gapu >= 0.0f) break;
      }
      if (j < a) continue;
      const Veh* u = &in->veh[j];
      int masku = sc->route_mask[sc->route_off[u->route] + u->cursor];
      if ((masku >> sc->lane_index[nl]) & 1) continue;                       /* its lane leads on: not urgent */
      if (!(sc->lane_len[nl] - u->pos < 60.0f || u->wait > 3.0f)) continue;
      int du = strategic_dir(sc, u, nl);
      if ((du > 0 ? sc->lane_left[nl] : (du < 0 ? sc->lane_right[nl] : -1)) != lane) continue;
      if (!(sc->lane_perm[lane] & sc->vtype_bit[u->vtype])) continue;
      vsafe = fminf(vsafe, follow_speed(gapu - COOP_MARGIN, u->speed, VT(s, u->vtype, VT_DECEL), decel, tau));
    }
  }
  float vmin_n = fmaxf(0.0f, v - decel);
  float vmin_e = fmaxf(0.0f, v - fmaxf(decel, EMERGENCY_DECEL));
  float vmin = fminf(vmin_n, fmaxf(vsafe, vmin_e));
  float vcand = fmaxf(vmin, vsafe);
  float vn = vcand;
  if (sigma > 0.0f) {
    uint32_t r[4];
    rng4(s, in, STREAM_DAWDLE, (uint32_t)x->vid, (uint32_t)in->tick, r);
    float xi = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
    vn = fmaxf(vmin, dawdle(vcand, accel, sigma, xi));
  }
  /* ---- lane-change decision (multi-lane normal edges) ---- */
  int target = -1;
  if (sc->lane_change && !sc->lane_internal[lane] && (sc->lane_left[lane] >= 0 || sc->lane_right[lane] >= 0)
      && x->pos + vn <= sc->lane_len[lane]) {
    int ro = sc->route_off[x->route];
    int mask = sc->route_mask[ro + x->cursor];
    int bestm = (mask >> 8) & 0xFF;
    int okm = mask & 0xFF, myidx = sc->lane_index[lane];
    int dir = strategic_dir(sc, x, lane);
    /* strategic (must): the route cannot continue from this lane.  A lane that leads on but is not "best"
     * only makes the best lanes attractive (no speed loss needed to go there); any lane that leads on may be
     * used to get around a blocked leader. */
    int strategic = dir != 0 && !((okm >> myidx) & 1);
    int cur_best = (bestm >> myidx) & 1;
    /* urgent: the route cannot continue from this lane and the lane end is near (or the vehicle already
     * stands): accept any gap the neighbours can still handle with emergency braking */
    int urgent = strategic && (sc->lane_len[lane] - x->pos < 60.0f || x->wait > 3.0f);
    for (int pass = 0; pass < 2; ++pass) {
      /* pass 0: strategic direction (if any); otherwise both directions, left then right */
      int d;
      if (strategic) { if (pass) break; d = dir; }
      else {
        if (x->lcc > 0 || (cur_best && !(vlead_limit < vacc - 1.0f))) break;
        d = pass == 0 ? 1 : -1;
      }
      if (d == 0) break;
      if ((d > 0) == ((in->tick & 1) != 0)) continue; /* even ticks: leftward, odd ticks: rightward */
      int nl = d > 0 ? sc->lane_left[lane] : sc->lane_right[lane];
      if (nl < 0 || !(sc->lane_perm[nl] & sc->vtype_bit[vt])) continue;
      if (!strategic && !((okm >> sc->lane_index[nl]) & 1)) continue;
      /* neighbours in nl: leader = last with pos >= mine, follower = first with pos < mine */
      int a = in->lane_start[nl], b = in->lane_start[nl + 1], j = a;
      while (j < b && in->veh[j].pos >= x->pos) ++j;
      float vfol = vacc;
      int ok = 1;
      if (j > a) {
        const Veh* ld = &in->veh[j - 1];
        float gap = ld->pos - VT(s, ld->vtype, VT_LEN) - x->pos - mingap;
        if (gap < 0.0f) ok = 0;
        else {
          vfol = follow_speed(gap, ld->speed, VT(s, ld->vtype, VT_DECEL), decel, tau);
          if (vfol < v - (urgent ? fmaxf(decel, EMERGENCY_DECEL) : decel)) ok = 0;
        }
      }
      if (ok && j < b) {
        const Veh* fo = &in->veh[j];
        float gap = x->pos - len - fo->pos - VT(s, fo->vtype, VT_GAP);
        if (gap < 0.0f) ok = 0;
        else if (urgent) {
          if (gap < brake_gap(fo->speed, fmaxf(VT(s, fo->vtype, VT_DECEL), EMERGENCY_DECEL), 1.0f)) ok = 0;   /* 1 s: it reacts a tick late */
        } else {
          float vf = follow_speed(gap, v, decel, VT(s, fo->vtype, VT_DECEL), VT(s, fo->vtype, VT_TAU));
          if (vf < fo->speed + VT(s, fo->vtype, VT_ACCEL) - VT(s, fo->vtype, VT_DECEL)) ok = 0;
        }
      }
      /* nobody behind in the target lane: the follower may still be upstream of it, about to come out of a junction */
      if (ok && j >= b && x->pos - len < 60.0f) {
        for (int w = sc->lane_watch_off[nl]; ok && w < sc->lane_watch_off[nl + 1]; ++w) {
          int pl = sc->lane_watch_lane[w];
          if (lane_count(in, pl) == 0) continue;
          const Veh* h = &in->veh[in->lane_start[pl]];
          float gap = (x->pos - len) + sc->lane_watch_dist[w] + (sc->lane_len[pl] - h->pos) - VT(s, h->vtype, VT_GAP);
          int unsafe;
          if (urgent) unsafe = gap < brake_gap(h->speed, fmaxf(VT(s, h->vtype, VT_DECEL), EMERGENCY_DECEL), 1.0f);
          else {
            float vf = follow_speed(gap, v, decel, VT(s, h->vtype, VT_DECEL), VT(s, h->vtype, VT_TAU));
            unsafe = vf < h->speed + VT(s, h->vtype, VT_ACCEL) - VT(s, h->vtype, VT_DECEL);
          }
          if (!unsafe) continue;
          int cur = pl, cc = h->cursor;          /* does its route lead onto the target lane? */
          for (int hop = 0; hop < 4; ++hop) {
            int k = choose_link(sc, cur, h->route, cc);
            if (k < 0) break;
            int nxt = sc->link_via[k] >= 0 ? sc->link_via[k] : sc->link_to[k];
            if (nxt == nl) { ok = 0; break; }
            if (!sc->lane_internal[nxt]) cc += 1;
            cur = nxt;
          }
        }
      }
      if (!ok) continue;
      if (!strategic) {   /* required speed gain: none towards a best lane, 1 m/s between best lanes, 2 m/s away from them */
        int nl_best = (bestm >> sc->lane_index[nl]) & 1;
        float gain = fminf(vfol, vacc) - vlead_limit;
        if (nl_best && !cur_best) { if (!(gain >= -0.5f)) continue; }
        else if (nl_best) { if (!(gain > 1.0f)) continue; }
        else if (!(gain > 2.0f && vlead_limit < vacc - 1.0f)) continue;
      }
      target = nl;
      vn = fmaxf(0.0f, fminf(vn, vfol));
      break;
    }
  }
  /* ---- deadlock breaker: two standing lane heads that each need the other's lane trade places ---- */
  if (target < 0 && wrong_lane_head && sc->lane_change && v < HALT_SPEED && sc->lane_len[lane] - x->pos < 1.0f) {
    int d = strategic_dir(sc, x, lane);
    int nl = d > 0 ? sc->lane_left[lane] : (d < 0 ? sc->lane_right[lane] : -1);
    if (nl >= 0 && lane_count(in, nl) > 0 && (sc->lane_perm[nl] & sc->vtype_bit[vt])) {
      const Veh* y = &in->veh[in->lane_start[nl]];
      if (y->speed < HALT_SPEED && sc->lane_len[nl] - y->pos < 1.0f && (sc->lane_perm[lane] & sc->vtype_bit[y->vtype])
          && choose_link(sc, nl, y->route, y->cursor) == -2 && strategic_dir(sc, y, nl) == -d) {
        target = nl;
        vn = 0.0f;
      }
    }
  }
  in->vnext[i] = vn;
  in->new_lane[i] = target >= 0 ? target : lane;
}

/* insertion safety against the heads of the lanes upstream of origin `o` (within the compiled reach): nobody who is
 * about to drive onto `lane` may be forced into hard braking by a vehicle standing with its back `back` metres in */
static int upstream_clear(const OrcSim* s, const Inst* in, int o, int lane, float back) {
  const RsScenario* sc = &s->sc;
  for (int w = sc->origin_watch_off[o]; w < sc->origin_watch_off[o + 1]; ++w) {
    int pl = sc->origin_watch_lane[w];
    if (in->lane_start2[pl + 1] <= in->lane_start2[pl]) continue;
    const Veh* h = &in->veh2[in->lane_start2[pl]];
    int cur = pl, cc = h->cursor, reaches = 0;
    for (int hop = 0; hop < 4; ++hop) {
      int k = choose_link(sc, cur, h->route, cc);
      if (k < 0) break;
      int nxt = sc->link_via[k] >= 0 ? sc->link_via[k] : sc->link_to[k];
      if (nxt == lane) { reaches = 1; break; }
      if (!sc->lane_internal[nxt]) cc += 1;
      cur = nxt;
    }
    if (!reaches) continue;
    float gap = ((sc->lane_len[pl] - h->pos) + sc->origin_watch_dist[w] - VT(s, h->vtype, VT_GAP)) + back;
    if (gap < brake_gap(h->speed, VT(s, h->vtype, VT_DECEL), VT(s, h->vtype, VT_TAU))) return 0;
  }
  return 1;
}

static int cmp_mover(const Inst* in, int a, int b) {
  float xa = in->new_pos[a], xb = in->new_pos[b];
  if (xa > xb) return -1;
  if (xa < xb) return 1;
  return a - b;
}

static void record_arrival(Inst* in, const Veh* x) {
  if (in->trip_arrival) {
    in->trip_arrival[x->vid] = in->tick; in->trip_depart_tick[x->vid] = x->depart;
    in->trip_tloss[x->vid] = x->tloss; in->trip_ddelay[x->vid] = x->ddelay; in->trip_wait[x->vid] = x->await;
  }
  in->st.n_arrived += 1;
  in->st.sum_delay_arrived += x->tloss + (float)x->ddelay;
  in->st.sum_duration_arrived += (float)(in->tick - x->depart);
  in->st.sum_wait_arrived += x->await;
}

static void tick_instance(OrcSim* s, Inst* in) {
  const RsScenario* sc = &s->sc;
  int L = sc->n_lanes;
  /* 0. traffic lights: static-program countdown */
  for (int t = 0; t < sc->n_tls; ++t) {
    int np = sc->tls_phase_off[t + 1] - sc->tls_phase_off[t];
    int guard = 0;
    while (in->tick >= in->tls_end[t] && guard++ < 64) {
      in->tls_phase[t] = (in->tls_phase[t] + 1) % np;
      int d = sc->phase_dur[sc->tls_phase_off[t] + in->tls_phase[t]];
      in->tls_end[t] += d > 0 ? d : 1;
    }
  }
  /* 1. lane tails */
  for (int l = 0; l < L; ++l) {
    int b = in->lane_start[l + 1];
    if (b > in->lane_start[l]) {
      const Veh* x = &in->veh[b - 1];
      in->tail_back[l] = x->pos - VT(s, x->vtype, VT_LEN);
      in->tail_speed[l] = x->speed;
      in->tail_decel[l] = VT(s, x->vtype, VT_DECEL);
    }
    float occ = 0.0f;
    for (int i = in->lane_start[l]; i < b; ++i) occ += VT(s, in->veh[i].vtype, VT_LEN) + VT(s, in->veh[i].vtype, VT_GAP);
    in->lane_occ[l] = occ;
  }
  /* 2. plan */
  for (int l = 0; l < L; ++l)
    for (int i = in->lane_start[l]; i < in->lane_start[l + 1]; ++i) plan_vehicle(s, in, i, l, i - in->lane_start[l]);
  /* 3. move */
  for (int l = 0; l < L; ++l) {
    for (int i = in->lane_start[l]; i < in->lane_start[l + 1]; ++i) {
      Veh* x = &in->veh[i];
      float vn = in->vnext[i];
      float vmaxl = fminf(sc->lane_vmax[l] * x->sf, VT(s, x->vtype, VT_VMAX));
      x->accel = vn - x->speed;
      x->speed = vn;
      x->wait = vn < HALT_SPEED ? x->wait + 1.0f : 0.0f;
      if (vn < HALT_SPEED) x->await += 1.0f;
      x->tloss += (vmaxl - vn) / vmaxl;
      if (x->lcc > 0) x->lcc -= 1;
      float p = x->pos + vn;
      int cur = l, cc = x->cursor;
      if (in->new_lane[i] != l) { cur = in->new_lane[i]; x->lcc = LC_COOLDOWN; } /* lateral move */
      else {
        int guard = 0;
        while (p > sc->lane_len[cur] && guard++ < 64) {
          int k = choose_link(sc, cur, x->route, cc);
          if (k == -1) { cur = -1; break; }
          if (k == -2) { p = sc->lane_len[cur]; break; } /* cannot happen: planned to stop */
          p -= sc->lane_len[cur];
          cur = sc->link_via[k] >= 0 ? sc->link_via[k] : sc->link_to[k];
          if (!sc->lane_internal[cur]) cc += 1;
        }
      }
      in->new_lane[i] = cur;
      in->new_pos[i] = p;
      in->new_cursor[i] = cc;
      if (cur < 0) record_arrival(in, x);
    }
  }
  /* 4/5. rebuild the lane-major order: stayers keep their order, movers merge in by position */
  {
    int n = in->n_veh;
    int* movers = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* lane_movers_off = (int*)calloc((size_t)L + 2, sizeof(int));
    int nm = 0;
    for (int l = 0; l < L; ++l)
      for (int i = in->lane_start[l]; i < in->lane_start[l + 1]; ++i)
        if (in->new_lane[i] >= 0 && in->new_lane[i] != l) lane_movers_off[in->new_lane[i] + 1] += 1;
    for (int l = 0; l < L; ++l) lane_movers_off[l + 1] += lane_movers_off[l];
    int* fill = (int*)calloc((size_t)L + 1, sizeof(int));
    for (int l = 0; l < L; ++l)
      for (int i = in->lane_start[l]; i < in->lane_start[l + 1]; ++i) {
        int nlane = in->new_lane[i];
        if (nlane >= 0 && nlane != l) { movers[lane_movers_off[nlane] + fill[nlane]++] = i; nm++; }
      }
    int w = 0;
    for (int l = 0; l < L; ++l) {
      in->lane_start2[l] = w;
      int m0 = lane_movers_off[l], m1 = lane_movers_off[l + 1];
      for (int u = m0 + 1; u < m1; ++u) { /* insertion sort: (new_pos desc, old index asc) */
        int key = movers[u], w2 = u - 1;
        while (w2 >= m0 && cmp_mover(in, movers[w2], key) > 0) { movers[w2 + 1] = movers[w2]; --w2; }
        movers[w2 + 1] = key;
      }
      int mi = m0;
      for (int i = in->lane_start[l]; i < in->lane_start[l + 1]; ++i) {
        if (in->new_lane[i] != l) continue; /* left the lane */
        while (mi < m1 && in->new_pos[movers[mi]] > in->new_pos[i]) {
          int j = movers[mi++];
          in->veh2[w] = in->veh[j]; in->veh2[w].pos = in->new_pos[j]; in->veh2[w].cursor = in->new_cursor[j]; ++w;
        }
        in->veh2[w] = in->veh[i]; in->veh2[w].pos = in->new_pos[i]; in->veh2[w].cursor = in->new_cursor[i]; ++w;
      }
      while (mi < m1) {
        int j = movers[mi++];
        in->veh2[w] = in->veh[j]; in->veh2[w].pos = in->new_pos[j]; in->veh2[w].cursor = in->new_cursor[j]; ++w;
      }
      /* 6. insertion at the back of origin lanes happens below, after all lanes are placed */
    }
    in->lane_start2[L] = w;
    free(movers); free(lane_movers_off); free(fill);
    (void)nm;
  }
  /* 6. insertion (departPos="base", departSpeed=0), decided on the post-move state */
  {
    int L1 = L + 1;
    int* add = (int*)calloc((size_t)L1, sizeof(int));
    Veh* newv = (Veh*)malloc(sizeof(Veh) * (size_t)(sc->n_origins > 0 ? sc->n_origins : 1));
    int* newl = (int*)malloc(sizeof(int) * (size_t)(sc->n_origins > 0 ? sc->n_origins : 1));
    int* newr = (int*)malloc(sizeof(int) * (size_t)(sc->n_origins > 0 ? sc->n_origins : 1));
    int nn = 0;
    int n_after_move = in->lane_start2[L];
    for (int o = 0; o < sc->n_origins; ++o) {
      int lane = sc->origin_lane[o];
      int route, vt, vid, ddelay;
      if (sc->synthetic) {
        uint32_t r[4];
        rng4(s, in, STREAM_DEMAND, (uint32_t)o, (uint32_t)in->tick, r);
        if ((int32_t)(r[0] >> 8) < sc->origin_rate[o]) in->origin_backlog[o] += 1;
        if (in->origin_backlog[o] <= 0) continue;
        int nr = sc->origin_route_off[o + 1] - sc->origin_route_off[o];
        if (nr <= 0) { in->origin_backlog[o] = 0; continue; }
        rng4(s, in, STREAM_ROUTE, (uint32_t)o, (uint32_t)in->origin_cur[o], r);
        route = sc->origin_route[sc->origin_route_off[o] + (int)(r[0] % (uint32_t)nr)];
        vt = sc->synthetic_vtype;
        vid = (o << 16) | (in->origin_cur[o] & 0xFFFF);
        ddelay = 0; /* aggregate depart delay is accumulated from the backlog below */
      } else {
        int c = sc->origin_off[o] + in->origin_cur[o];
        if (c >= sc->origin_off[o + 1]) continue;
        if (sc->trip_depart[c] > (float)in->tick) continue;
        route = sc->trip_route[c]; vt = sc->trip_vtype[c]; vid = c;
        ddelay = in->tick - (int)sc->trip_depart[c];
      }
      float len = VT(s, vt, VT_LEN), mingap = VT(s, vt, VT_GAP);
      int a = in->lane_start2[lane], b = in->lane_start2[lane + 1];
      int ok = 0, rank = -1;        /* rank: vehicles of the lane ahead of the newcomer; -1 = it goes to the back */
      float ipos = len;
      /* departPos="random_free" (arterial4x4's route files): SUMO tries ten uniformly drawn positions for one where the
       * vehicle fits (MSLane::insertVehicle, RANDOM_FREE), then falls back to a free insertion -- here the base rule.
       * A position fits when no vehicle of the lane ahead of it is closer than the minimum gap and none behind it
       * (on the lane, else the upstream heads) would have to brake hard for a standing vehicle at that place. */
      if (!sc->synthetic && sc->trip_depart_pos[vid] == 1 && len <= sc->lane_len[lane]) {
        for (int k = 0; k < 10 && !ok; ++k) {
          uint32_t r[4];
          rng4(s, in, STREAM_DEPARTPOS, (uint32_t)vid, (uint32_t)in->tick * 4u + (uint32_t)(k >> 2), r);
          float u = (float)(r[k & 3] >> 8) * (1.0f / 16777216.0f);
          float p = len + u * (sc->lane_len[lane] - len);
          float back = p - len;
          int fits = 1, ahead = 0, behind = 0;
          for (int i = a; i < b; ++i) {
            const Veh* x = &in->veh2[i];
            if (x->pos >= p) {
              ahead += 1;
              if (x->pos - VT(s, x->vtype, VT_LEN) - p - mingap < 0.0f) fits = 0;
            } else {
              behind += 1;
              if (back - x->pos - VT(s, x->vtype, VT_GAP) < brake_gap(x->speed, VT(s, x->vtype, VT_DECEL), VT(s, x->vtype, VT_TAU))) fits = 0;
            }
          }
          if (fits && behind == 0) fits = upstream_clear(s, in, o, lane, back);
          if (fits) { ok = 1; rank = ahead; ipos = p; }
        }
      }
      if (!ok) {
        ok = len <= sc->lane_len[lane];
        if (ok && b > a) {
          const Veh* tl = &in->veh2[b - 1];
          if (tl->pos - VT(s, tl->vtype, VT_LEN) - len - mingap < 0.0f) ok = 0;
        }
        /* upstream safety: nobody who is about to drive onto this lane may be forced into hard braking */
        if (ok) ok = upstream_clear(s, in, o, lane, 0.0f);
      }
      /* capacity of the vehicle store: an insertion the road has room for but the store does not is put off and COUNTED
       * (SUMO has no such limit; a non-zero count means the run was truncated by `vcap`) */
      if (ok && !((n_after_move + nn) < sc->vcap)) { ok = 0; in->st.n_cap_refused += 1; }
      if (ok) {
        Veh nv; memset(&nv, 0, sizeof nv);
        nv.pos = ipos; nv.speed = 0.0f; nv.vid = vid; nv.vtype = vt; nv.route = route; nv.cursor = 0;
        nv.depart = in->tick; nv.ddelay = ddelay; nv.seen_epoch = -2; nv.seen_sig = -1;
        float dev = sc->speed_dev_override >= 0.0f ? sc->speed_dev_override : VT(s, vt, VT_DEV);
        nv.sf = speed_factor(s, in, vid, dev);
        newv[nn] = nv; newl[nn] = lane; newr[nn] = rank; add[lane] += 1; ++nn;
        in->origin_cur[o] += 1;
        if (sc->synthetic) in->origin_backlog[o] -= 1;
        in->st.n_inserted += 1;
      }
    }
    /* final CSR: veh2 + newcomers appended at each origin lane's back */
    int w = 0;
    for (int l = 0; l < L; ++l) {
      int a = in->lane_start2[l], b = in->lane_start2[l + 1];
      in->lane_start[l] = w;
      int at = -1, nq = -1;      /* one origin per lane: at most one newcomer, at its rank or at the back */
      if (add[l])
        for (int q = 0; q < nn; ++q) if (newl[q] == l) { nq = q; at = newr[q] < 0 ? b : a + newr[q]; }
      for (int i = a; i < b; ++i) {
        if (i == at) in->veh[w++] = newv[nq];
        in->veh[w++] = in->veh2[i];
      }
      if (at == b) in->veh[w++] = newv[nq];
    }
    in->lane_start[L] = w;
    in->n_veh = w;
    free(add); free(newv); free(newl); free(newr);
  }
  if (getenv("ORC_TRACE")) {   /* diagnostic: ORC_TRACE="vidA,vidB" prints both vehicles every tick */
    int va = -1, vb = -1;
    sscanf(getenv("ORC_TRACE"), "%d,%d", &va, &vb);
    for (int l = 0; l < L; ++l)
      for (int i = in->lane_start[l]; i < in->lane_start[l + 1]; ++i)
        if (in->veh[i].vid == va || in->veh[i].vid == vb)
          fprintf(stderr, "[trace] env %llu tick %d vid %d lane %d pos %.3f v %.3f wait %.0f\n", (unsigned long long)in->env_id,
                  in->tick, in->veh[i].vid, l, in->veh[i].pos, in->veh[i].speed, in->veh[i].wait);
  }
  /* ordering anomaly check (diagnostic; stays 0 when the rules keep vehicles apart) */
  for (int l = 0; l < L; ++l)
    for (int i = in->lane_start[l] + 1; i < in->lane_start[l + 1]; ++i)
      if (in->veh[i].pos > in->veh[i - 1].pos) {
        in->st.anomalies += 1;
        if (getenv("ORC_DEBUG"))
          fprintf(stderr, "[oracle] ordering anomaly: env %llu tick %d lane %d: vid %d pos %.3f v %.3f behind vid %d pos %.3f v %.3f\n",
                  (unsigned long long)in->env_id, in->tick, l, in->veh[i].vid, in->veh[i].pos, in->veh[i].speed,
                  in->veh[i - 1].vid, in->veh[i - 1].pos, in->veh[i - 1].speed);
      }
  if (sc->synthetic)
    for (int o = 0; o < sc->n_origins; ++o) in->st.sum_delay_pending += (float)in->origin_backlog[o];
  in->st.sum_active_ticks += in->n_veh;
  in->tick += 1;
}

/* ------------------------------------------------------------------ RESCO layer */
static void set_phase(OrcSim* s, Inst* in, int sig, int idx) {
  const RsScenario* sc = &s->sc;
  int t = sc->sig_tls[sig];
  int np = sc->tls_phase_off[t + 1] - sc->tls_phase_off[t];
  if (idx < 0 || idx >= np) return;
  in->tls_phase[t] = idx;
  in->tls_end[t] = in->tick + sc->phase_dur[sc->tls_phase_off[t] + idx];
}

/* Signal.prep_phase, traffic_signal.py:176-184 */
static void prep_phase(OrcSim* s, Inst* in, int sig, int act) {
  const RsScenario* sc = &s->sc;
  int cur = in->tls_phase[sc->sig_tls[sig]];
  if (cur == act) { in->next_phase[sig] = cur; return; }
  in->next_phase[sig] = act;
  int ng = sc->sig_n_green[sig];
  if (cur >= 0 && cur < ng && act >= 0 && act < ng) {
    int y = sc->yellow_idx[sc->sig_yellow_off[sig] + cur * ng + act];
    if (y >= 0) set_phase(s, in, sig, y);
  }
}

/* Signal.observe (traffic_signal.py:189-235) + states.mplight/wave + rewards.* + calc_metrics */
static void observe_instance(OrcSim* s, Inst* in) {
  const RsScenario* sc = &s->sc;
  int e = in->epoch;
  for (int sg = 0; sg < sc->n_signals; ++sg) {
    for (int q = sc->sig_lane_off[sg]; q < sc->sig_lane_off[sg + 1]; ++q) {
      int lane = sc->sig_lane[q];
      float queue = 0, appr = 0, tw = 0, mw = 0, ss = 0, arrv = 0;
      float tdist = sc->lane_tls_dist[lane];
      for (int i = in->lane_start[lane]; i < in->lane_start[lane + 1]; ++i) {
        Veh* x = &in->veh[i];
        if (tdist < 0.0f) continue;                                  /* len(path) == 0 */
        float dist = (sc->lane_len[lane] - x->pos) + tdist;
        if (!(dist <= sc->max_distance)) continue;                   /* detector range */
        int contiguous = (x->seen_epoch == e - 1 && x->seen_sig == sg);
        if (!contiguous) { x->rwait = 0.0f; arrv += 1.0f; }          /* popped on departure / never seen: an arrival */
        if (x->rwait > 0.0f) x->rwait += (float)sc->step_length;     /* `vehicle in self.waiting_times` */
        else if (x->wait > 0.0f) x->rwait = x->wait;                 /* getWaitingTime() > 0 */
        x->seen_epoch = e; x->seen_sig = sg;
        if (x->rwait > 0.0f) { tw += x->rwait; queue += 1.0f; if (x->rwait > mw) mw = x->rwait; }
        else appr += 1.0f;
        ss += x->speed;
      }
      in->lane_queue[q] = queue; in->lane_approach[q] = appr; in->lane_total_wait[q] = tw;
      in->lane_max_wait[q] = mw; in->lane_speed_sum[q] = ss; in->lane_arrivals[q] = arrv;
    }
  }
  in->epoch += 1;
  for (int sg = 0; sg < sc->n_signals; ++sg) {
    int q0 = sc->sig_lane_off[sg], q1 = sc->sig_lane_off[sg + 1];
    in->phase_obs[sg] = in->tls_phase[sc->sig_tls[sg]];
    float* mp = in->mplight + sg * 13;
    float* wv = in->wave + sg * 12;
    mp[0] = (float)in->phase_obs[sg];
    for (int m = 0; m < 12; ++m) {
      float qsum = 0, wsum = 0;
      for (int j = sc->mv_off[sg * 12 + m]; j < sc->mv_off[sg * 12 + m + 1]; ++j) {
        int q = q0 + sc->mv_lane[j];
        qsum += in->lane_queue[q];
        wsum += in->lane_queue[q] + in->lane_approach[q];
      }
      for (int j = sc->mvo_off[sg * 12 + m]; j < sc->mvo_off[sg * 12 + m + 1]; ++j)
        qsum -= in->lane_queue[sc->sig_lane_off[sc->mvo_sig[j]] + sc->mvo_slot[j]];
      mp[1 + m] = qsum; wv[m] = wsum;
    }
    float tw = 0, ql = 0, mq = 0;
    for (int q = q0; q < q1; ++q) {
      tw += in->lane_total_wait[q]; ql += in->lane_queue[q];
      if (in->lane_queue[q] > mq) mq = in->lane_queue[q];
    }
    /* states.drq / drq_norm (states.py:6-59): the one-hot compares the LANE index with the phase index */
    for (int q = q0; q < q1; ++q) {
      float oh = (q - q0) == in->phase_obs[sg] ? 1.0f : 0.0f;
      float* d = in->drq + (size_t)q * 5; float* dn = in->drq_norm + (size_t)q * 5;
      d[0] = oh; d[1] = in->lane_approach[q]; d[2] = in->lane_total_wait[q]; d[3] = in->lane_queue[q]; d[4] = in->lane_speed_sum[q];
      dn[0] = oh; dn[1] = in->lane_approach[q] / 28.0f; dn[2] = in->lane_total_wait[q] / 28.0f; dn[3] = in->lane_queue[q] / 28.0f;
      dn[4] = in->lane_speed_sum[q] / 20.0f / 28.0f;
    }
    /* states.mplight_full (states.py:83-113): total_speed is reset inside the lane loop -> last lane's speed sum */
    {
      float* mf = in->mplight_full + (size_t)sg * 49;
      mf[0] = (float)in->phase_obs[sg];
      for (int m = 0; m < 12; ++m) {
        float wsum = 0, spd = 0, asum = 0;
        for (int j = sc->mv_off[sg * 12 + m]; j < sc->mv_off[sg * 12 + m + 1]; ++j) {
          int q = q0 + sc->mv_lane[j];
          wsum += in->lane_total_wait[q] / 28.0f; spd = in->lane_speed_sum[q]; asum += in->lane_approach[q] / 28.0f;
        }
        mf[1 + 4 * m] = mp[1 + m]; mf[2 + 4 * m] = wsum; mf[3 + 4 * m] = spd; mf[4 + 4 * m] = asum;
      }
    }
    in->rew_wait[sg] = -tw;
    in->rew_wait_norm[sg] = fminf(fmaxf(-tw / 224.0f, -4.0f), 4.0f);
    float pr = ql;
    for (int j = sc->out_off[sg]; j < sc->out_off[sg + 1]; ++j)
      pr -= in->lane_queue[sc->sig_lane_off[sc->out_sig[j]] + sc->out_slot[j]];
    in->rew_pressure[sg] = -pr;
    in->sig_queue_len[sg] = (int32_t)ql; in->sig_max_queue[sg] = (int32_t)mq;
  }
}

/* ------------------------------------------------------------------ public oracle API */
static void* own(OrcSim* s, size_t bytes) {
  void* p = calloc(1, bytes ? bytes : 1);
  s->owned = (void**)realloc(s->owned, sizeof(void*) * (size_t)(s->n_owned + 1));
  s->owned[s->n_owned++] = p;
  return p;
}
#define DUP(field, count, type) do { size_t _b = sizeof(type) * (size_t)(count); void* _p = own(s, _b); \
  if (sc->field && _b) { memcpy(_p, sc->field, _b); } \
  s->sc.field = (const type*)_p; } while (0)

OrcSim* orc_create(const RsScenario* sc, int32_t n_env, uint64_t seed) {
  if (!sc || sc->abi_version != RS_ABI_VERSION || n_env <= 0) return NULL;
  OrcSim* s = (OrcSim*)calloc(1, sizeof(OrcSim));
  s->sc = *sc; s->n_env = n_env; s->seed = seed;
  int L = sc->n_lanes, K = sc->n_links, S = sc->n_signals;
  DUP(lane_len, L, float); DUP(lane_vmax, L, float); DUP(lane_edge, L, int32_t); DUP(lane_index, L, int32_t);
  DUP(lane_perm, L, int32_t); DUP(lane_internal, L, int32_t); DUP(lane_left, L, int32_t); DUP(lane_right, L, int32_t);
  DUP(lane_link_off, L + 1, int32_t); DUP(lane_tls_dist, L, float); DUP(lane_sig, L, int32_t); DUP(lane_sig_slot, L, int32_t);
  DUP(edge_lane0, sc->n_edges, int32_t); DUP(edge_nlanes, sc->n_edges, int32_t);
  DUP(link_from, K, int32_t); DUP(link_to, K, int32_t); DUP(link_via, K, int32_t); DUP(link_tls, K, int32_t);
  DUP(link_tlidx, K, int32_t); DUP(link_state, K, int32_t); DUP(link_to_edge, K, int32_t); DUP(link_via_len, K, float);
  DUP(link_last_int, K, int32_t); DUP(link_cont, K, int32_t); DUP(link_parent, K, int32_t); DUP(link_foe_off, K + 1, int32_t);
  DUP(foe_link, sc->n_foes, int32_t); DUP(foe_flags, sc->n_foes, int32_t);
  DUP(tls_phase_off, sc->n_tls + 1, int32_t); DUP(tls_nlinks, sc->n_tls, int32_t); DUP(tls_init_phase, sc->n_tls, int32_t);
  DUP(tls_init_left, sc->n_tls, int32_t); DUP(phase_dur, sc->n_phases, int32_t); DUP(phase_state_off, sc->n_phases, int32_t);
  DUP(state_chars, sc->n_state_chars, uint8_t);
  DUP(sig_tls, S, int32_t); DUP(sig_n_green, S, int32_t); DUP(sig_yellow_off, S + 1, int32_t); DUP(yellow_idx, sc->n_yellow, int32_t);
  DUP(sig_lane_off, S + 1, int32_t); DUP(sig_lane, sc->n_sig_lanes, int32_t);
  DUP(mv_off, S * 12 + 1, int32_t); DUP(mv_lane, sc->n_mv_lanes, int32_t);
  DUP(mvo_off, S * 12 + 1, int32_t); DUP(mvo_sig, sc->n_mvo, int32_t); DUP(mvo_slot, sc->n_mvo, int32_t);
  DUP(out_off, S + 1, int32_t); DUP(out_sig, sc->n_out, int32_t); DUP(out_slot, sc->n_out, int32_t);
  DUP(vtype, sc->n_vtypes * 8, float); DUP(vtype_bit, sc->n_vtypes, int32_t);
  DUP(route_off, sc->n_routes + 1, int32_t); DUP(route_edge, sc->n_route_steps, int32_t); DUP(route_mask, sc->n_route_steps, int32_t);
  DUP(origin_lane, sc->n_origins, int32_t); DUP(origin_off, sc->n_origins + 1, int32_t);
  DUP(trip_depart, sc->n_trips, float); DUP(trip_route, sc->n_trips, int32_t); DUP(trip_vtype, sc->n_trips, int32_t);
  DUP(trip_file, sc->n_trips, int32_t); DUP(trip_depart_pos, sc->n_trips, int32_t);
  DUP(origin_rate, sc->n_origins, int32_t); DUP(origin_route_off, sc->n_origins + 1, int32_t);
  DUP(origin_route, sc->n_origin_routes, int32_t);
  DUP(origin_watch_off, sc->n_origins + 1, int32_t); DUP(origin_watch_lane, sc->n_watch, int32_t); DUP(origin_watch_dist, sc->n_watch, float); DUP(origin_watch_owner, sc->n_watch, int32_t);
  DUP(lane_watch_off, L + 1, int32_t); DUP(lane_watch_lane, sc->n_lane_watch, int32_t); DUP(lane_watch_dist, sc->n_lane_watch, float);
  s->inst = (Inst*)calloc((size_t)n_env, sizeof(Inst));
  int V = sc->vcap, SL = sc->n_sig_lanes;
  for (int e = 0; e < n_env; ++e) {
    Inst* in = &s->inst[e];
    in->veh = (Veh*)own(s, sizeof(Veh) * (size_t)V); in->veh2 = (Veh*)own(s, sizeof(Veh) * (size_t)V);
    in->lane_start = (int32_t*)own(s, 4 * (size_t)(L + 1)); in->lane_start2 = (int32_t*)own(s, 4 * (size_t)(L + 1));
    in->tls_phase = (int32_t*)own(s, 4 * (size_t)sc->n_tls); in->tls_end = (int32_t*)own(s, 4 * (size_t)sc->n_tls);
    in->next_phase = (int32_t*)own(s, 4 * (size_t)S);
    in->origin_cur = (int32_t*)own(s, 4 * (size_t)sc->n_origins); in->origin_backlog = (int32_t*)own(s, 4 * (size_t)sc->n_origins);
    in->tail_back = (float*)own(s, 4 * (size_t)L); in->tail_speed = (float*)own(s, 4 * (size_t)L); in->tail_decel = (float*)own(s, 4 * (size_t)L); in->lane_occ = (float*)own(s, 4 * (size_t)L);
    in->vnext = (float*)own(s, 4 * (size_t)V); in->new_lane = (int32_t*)own(s, 4 * (size_t)V);
    in->new_pos = (float*)own(s, 4 * (size_t)V); in->new_cursor = (int32_t*)own(s, 4 * (size_t)V);
    in->lane_queue = (float*)own(s, 4 * (size_t)SL); in->lane_approach = (float*)own(s, 4 * (size_t)SL);
    in->lane_total_wait = (float*)own(s, 4 * (size_t)SL); in->lane_max_wait = (float*)own(s, 4 * (size_t)SL);
    in->lane_speed_sum = (float*)own(s, 4 * (size_t)SL);
    in->lane_arrivals = (float*)own(s, 4 * (size_t)SL);
    in->phase_obs = (int32_t*)own(s, 4 * (size_t)S);
    in->mplight = (float*)own(s, 4 * (size_t)S * 13); in->wave = (float*)own(s, 4 * (size_t)S * 12);
    in->rew_wait = (float*)own(s, 4 * (size_t)S); in->rew_wait_norm = (float*)own(s, 4 * (size_t)S);
    in->rew_pressure = (float*)own(s, 4 * (size_t)S);
    in->drq = (float*)own(s, 4 * (size_t)SL * 5); in->drq_norm = (float*)own(s, 4 * (size_t)SL * 5);
    in->mplight_full = (float*)own(s, 4 * (size_t)S * 49);
    in->sig_queue_len = (int32_t*)own(s, 4 * (size_t)S); in->sig_max_queue = (int32_t*)own(s, 4 * (size_t)S);
    if (sc->record_trips && !sc->synthetic) {
      in->trip_arrival = (int32_t*)own(s, 4 * (size_t)sc->n_trips); in->trip_depart_tick = (int32_t*)own(s, 4 * (size_t)sc->n_trips);
      in->trip_ddelay = (int32_t*)own(s, 4 * (size_t)sc->n_trips); in->trip_tloss = (float*)own(s, 4 * (size_t)sc->n_trips);
      in->trip_wait = (float*)own(s, 4 * (size_t)sc->n_trips);
    }
  }
  return s;
}

void orc_destroy(OrcSim* s) {
  if (!s) return;
  for (int i = 0; i < s->n_owned; ++i) free(s->owned[i]);
  free(s->owned); free(s->inst); free(s);
}

void orc_reset(OrcSim* s, uint64_t seed, int64_t first_env_id) {
  const RsScenario* sc = &s->sc;
  s->seed = seed;
  for (int e = 0; e < s->n_env; ++e) {
    Inst* in = &s->inst[e];
    in->tick = 0; in->n_veh = 0; in->epoch = 0; in->env_id = (uint64_t)(first_env_id + e);
    memset(in->lane_start, 0, 4 * (size_t)(sc->n_lanes + 1));
    memset(in->origin_cur, 0, 4 * (size_t)sc->n_origins);
    memset(in->origin_backlog, 0, 4 * (size_t)sc->n_origins);
    memset(in->next_phase, 0, 4 * (size_t)sc->n_signals);
    memset(&in->st, 0, sizeof in->st);
    for (int t = 0; t < sc->n_tls; ++t) { in->tls_phase[t] = sc->tls_init_phase[t]; in->tls_end[t] = sc->tls_init_left[t]; }
    if (in->trip_arrival) for (int i = 0; i < sc->n_trips; ++i) in->trip_arrival[i] = -1;
  }
}

/* per-origin range of the trip table that is the next episode's demand (rs_set_demand_window) */
int orc_set_demand_window(OrcSim* s, const int32_t* origin_off) {
  if (s->sc.synthetic) return -1;
  memcpy((int32_t*)s->sc.origin_off, origin_off, sizeof(int32_t) * (size_t)(s->sc.n_origins + 1));
  return 0;
}

void orc_set_phase(OrcSim* s, const int32_t* phase, const uint8_t* mask) {
  int S = s->sc.n_signals;
  for (int e = 0; e < s->n_env; ++e)
    for (int sg = 0; sg < S; ++sg)
      if (!mask || mask[e * S + sg]) set_phase(s, &s->inst[e], sg, phase[e * S + sg]);
}

void orc_tick(OrcSim* s, int32_t n) {
  for (int e = 0; e < s->n_env; ++e)
    for (int i = 0; i < n; ++i) tick_instance(s, &s->inst[e]);
}

void orc_observe(OrcSim* s) {
  for (int e = 0; e < s->n_env; ++e) observe_instance(s, &s->inst[e]);
}

/* MultiSignal.step, multi_signal.py:164-197 */
void orc_env_step(OrcSim* s, const int32_t* actions) {
  const RsScenario* sc = &s->sc;
  int S = sc->n_signals;
  for (int e = 0; e < s->n_env; ++e) {
    Inst* in = &s->inst[e];
    for (int sg = 0; sg < S; ++sg) prep_phase(s, in, sg, actions[e * S + sg]);
    for (int i = 0; i < sc->yellow_length; ++i) tick_instance(s, in);
    for (int sg = 0; sg < S; ++sg) set_phase(s, in, sg, in->next_phase[sg]);
    for (int i = 0; i < sc->step_length - sc->yellow_length; ++i) tick_instance(s, in);
    observe_instance(s, in);
  }
}

/* host copies: each array [n_env, ...] */
void orc_get_obs(OrcSim* s, float* lane_queue, float* lane_approach, float* lane_total_wait, float* lane_max_wait,
                 float* lane_speed_sum, int32_t* phase, float* mplight, float* wave, float* rew_wait,
                 float* rew_wait_norm, float* rew_pressure, int32_t* sig_queue_len, int32_t* sig_max_queue,
                 float* lane_arrivals, float* drq, float* drq_norm, float* mplight_full) {
  int S = s->sc.n_signals, SL = s->sc.n_sig_lanes;
  for (int e = 0; e < s->n_env; ++e) {
    const Inst* in = &s->inst[e];
#define CP(dst, src, n, T) if (dst) memcpy((dst) + (size_t)e * (size_t)(n), (src), sizeof(T) * (size_t)(n))
    CP(lane_queue, in->lane_queue, SL, float); CP(lane_approach, in->lane_approach, SL, float);
    CP(lane_total_wait, in->lane_total_wait, SL, float); CP(lane_max_wait, in->lane_max_wait, SL, float);
    CP(lane_speed_sum, in->lane_speed_sum, SL, float); CP(phase, in->phase_obs, S, int32_t);
    CP(mplight, in->mplight, S * 13, float); CP(wave, in->wave, S * 12, float);
    CP(rew_wait, in->rew_wait, S, float); CP(rew_wait_norm, in->rew_wait_norm, S, float);
    CP(rew_pressure, in->rew_pressure, S, float);
    CP(sig_queue_len, in->sig_queue_len, S, int32_t); CP(sig_max_queue, in->sig_max_queue, S, int32_t);
    CP(lane_arrivals, in->lane_arrivals, SL, float);
    CP(drq, in->drq, SL * 5, float); CP(drq_norm, in->drq_norm, SL * 5, float); CP(mplight_full, in->mplight_full, S * 49, float);
#undef CP
  }
}

void orc_get_stats(OrcSim* s, RsStats* out) {
  const RsScenario* sc = &s->sc;
  for (int e = 0; e < s->n_env; ++e) {
    Inst* in = &s->inst[e];
    RsStats st = in->st;
    st.tick = in->tick; st.n_active = in->n_veh;
    float run = 0;
    for (int i = 0; i < in->n_veh; ++i) run += in->veh[i].tloss + (float)in->veh[i].ddelay;
    st.sum_delay_running = run;
    int backlog = 0;
    if (!sc->synthetic) {
      float pend = 0;
      for (int o = 0; o < sc->n_origins; ++o)
        for (int c = sc->origin_off[o] + in->origin_cur[o]; c < sc->origin_off[o + 1]; ++c)
          if (sc->trip_depart[c] <= (float)in->tick) { pend += (float)in->tick - sc->trip_depart[c]; backlog++; }
      st.sum_delay_pending = pend;
    } else {
      for (int o = 0; o < sc->n_origins; ++o) backlog += in->origin_backlog[o];
    }
    st.n_backlog = backlog;
    out[e] = st;
  }
}

int orc_dump_vehicles(OrcSim* s, int32_t env, int32_t* lane, float* pos, float* speed, float* accel, float* wait,
                      float* rwait, float* tloss, int32_t* vid, int32_t* vtype, int32_t* route, int32_t* cursor,
                      float* sf, int32_t* depart, float* acc_wait) {
  const Inst* in = &s->inst[env];
  for (int l = 0; l < s->sc.n_lanes; ++l)
    for (int i = in->lane_start[l]; i < in->lane_start[l + 1]; ++i) {
      const Veh* x = &in->veh[i];
      lane[i] = l; pos[i] = x->pos; speed[i] = x->speed; accel[i] = x->accel; wait[i] = x->wait; rwait[i] = x->rwait;
      tloss[i] = x->tloss; vid[i] = x->vid; vtype[i] = x->vtype; route[i] = x->route; cursor[i] = x->cursor;
      sf[i] = x->sf; depart[i] = x->depart; if (acc_wait) acc_wait[i] = x->await;
    }
  return in->n_veh;
}

int orc_get_trip_records(OrcSim* s, int32_t env, int32_t* arrival, int32_t* depart, float* tloss, int32_t* ddelay, float* wait) {
  const Inst* in = &s->inst[env];
  if (!in->trip_arrival) return -1;
  size_t n = (size_t)s->sc.n_trips;
  memcpy(arrival, in->trip_arrival, 4 * n); memcpy(depart, in->trip_depart_tick, 4 * n);
  memcpy(tloss, in->trip_tloss, 4 * n); memcpy(ddelay, in->trip_ddelay, 4 * n);
  if (wait) memcpy(wait, in->trip_wait, 4 * n);
  return 0;
}

void orc_get_phases(OrcSim* s, int32_t env, int32_t* tls_phase) {
  memcpy(tls_phase, s->inst[env].tls_phase, 4 * (size_t)s->sc.n_tls);
}

/* diagnostics: why is the head vehicle of `lane` not moving?  out = {link, reason, hop, seen*100} */
void orc_explain(OrcSim* s, int32_t env, int32_t lane, int32_t* out) {
  const RsScenario* sc = &s->sc;
  Inst* in = &s->inst[env];
  out[0] = out[1] = out[2] = out[3] = -9;
  if (lane_count(in, lane) == 0) return;
  for (int l = 0; l < sc->n_lanes; ++l) {
    int b = in->lane_start[l + 1];
    if (b > in->lane_start[l]) {
      const Veh* y = &in->veh[b - 1];
      in->tail_back[l] = y->pos - VT(s, y->vtype, VT_LEN); in->tail_speed[l] = y->speed; in->tail_decel[l] = VT(s, y->vtype, VT_DECEL);
    }
    float occ = 0.0f;
    for (int i = in->lane_start[l]; i < b; ++i) occ += VT(s, in->veh[i].vtype, VT_LEN) + VT(s, in->veh[i].vtype, VT_GAP);
    in->lane_occ[l] = occ;
  }
  const Veh* x = &in->veh[in->lane_start[lane]];
  float seen = sc->lane_len[lane] - x->pos;
  int cur = lane, cc = x->cursor;
  for (int hop = 0; hop < MAX_HOPS; ++hop) {
    int k = choose_link(sc, cur, x->route, cc);
    out[0] = k; out[2] = hop; out[3] = (int)(seen * 100.0f);
    if (k == -1) { out[1] = -1; return; }
    if (k == -2) { out[1] = 8; return; }
    int r = must_stop(s, in, x, k, seen, hop, cc, 1);
    if (r) { out[1] = r; return; }
    int nxt = sc->link_via[k] >= 0 ? sc->link_via[k] : sc->link_to[k];
    if (lane_count(in, nxt) > 0) { out[1] = 9; out[0] = nxt; return; }
    seen += sc->lane_len[nxt];
    if (!sc->lane_internal[nxt]) cc += 1;
    cur = nxt;
  }
  out[1] = 0;
}

/* scalar primitives exported for unit tests */
float orc_brake_gap(float v, float d, float h) { return brake_gap(v, d, h); }
float orc_max_safe_stop_speed(float g, float d, float t) { return max_safe_stop_speed(g, d, t); }
float orc_follow_speed(float g, float vl, float dl, float d, float t) { return follow_speed(g, vl, dl, d, t); }
float orc_free_speed(float d, float dist, float target) { return free_speed(d, dist, target); }
