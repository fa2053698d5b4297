/*
 * resco_b200.h -- C-ABI of the B200-native traffic-microsimulation backend.
 *
 * The reference (Pi-Star-Lab/RESCO) has no plugin API of its own; the seam it offers is the
 * `self.sumo` attribute (resco_benchmark/multi_signal.py:44,47,134,137; traffic_signal.py:29),
 * i.e. the TraCI/libsumo call surface listed in SURVEY.md §8(b).  Each entry point below names the
 * TraCI call(s) it replaces.  Plain pointers and sizes only; no torch / CUDA types in signatures
 * (streams travel as `void*`).  Every function returns 0 on success or a negative RsStatus, with a
 * message available from rs_last_error().  One host thread per RsSim.  All device work is enqueued
 * on the stream passed to the call; results are valid after that stream is synchronised (calls that
 * return data through HOST pointers synchronise themselves).
 */
#ifndef RESCO_B200_H
#define RESCO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_ABI_VERSION 5
#define RS_N_MOVEMENTS 12   /* the 12 movement keys of signal_config.py lane_sets */

typedef enum RsStatus {
  RS_OK = 0,
  RS_ERR_INVALID = -1,   /* bad argument / inconsistent scenario */
  RS_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
  RS_ERR_NOMEM = -3,
  RS_ERR_NODEVICE = -4,  /* no sm_100 device: the product path has NO CPU fallback */
  RS_ERR_CAPACITY = -5
} RsStatus;

/* Flat scenario tables (host pointers, copied to the device by rs_create).  Produced by
 * resco_b200/scenario/compiler.py from the SUMO net/route files the reference passes to
 * `traci.start` (multi_signal.py:117-137) and from config/signal_config.py. */
typedef struct RsScenario {
  int32_t abi_version;
  /* sizes */
  int32_t n_lanes, n_edges, n_links, n_foes, n_tls, n_phases, n_state_chars, n_signals;
  int32_t n_sig_lanes, n_mv_lanes, n_mvo, n_out, n_yellow;
  int32_t n_vtypes, n_routes, n_route_steps, n_origins, n_trips, n_origin_routes, n_watch, n_lane_watch;
  /* lanes */
  const float* lane_len;
  const float* lane_vmax;
  const int32_t* lane_edge;
  const int32_t* lane_index;
  const int32_t* lane_perm;
  const int32_t* lane_internal;
  const int32_t* lane_left;
  const int32_t* lane_right;
  const int32_t* lane_link_off;      /* [n_lanes+1] */
  const float* lane_tls_dist;        /* end-of-lane -> next TLS stop line, <0: none */
  const int32_t* lane_sig;           /* signal whose lane list holds the lane, or -1 */
  const int32_t* lane_sig_slot;      /* position in that signal's lane list */
  /* edges */
  const int32_t* edge_lane0;
  const int32_t* edge_nlanes;
  /* links (sorted by from-lane) */
  const int32_t* link_from;
  const int32_t* link_to;
  const int32_t* link_via;
  const int32_t* link_tls;
  const int32_t* link_tlidx;
  const int32_t* link_state;
  const int32_t* link_to_edge;
  const float* link_via_len;
  const int32_t* link_last_int;
  const int32_t* link_cont;
  const int32_t* link_parent;
  const int32_t* link_foe_off;       /* [n_links+1] */
  const int32_t* foe_link;
  const int32_t* foe_flags;          /* 1: must yield (response), 2: conflict (foes), 4: mutual */
  /* traffic-light programs as installed (Signal.__init__, traffic_signal.py:93-100) */
  const int32_t* tls_phase_off;      /* [n_tls+1] into phase_* */
  const int32_t* tls_nlinks;         /* state-string length */
  const int32_t* tls_init_phase;
  const int32_t* tls_init_left;      /* ticks left in the initial phase */
  const int32_t* phase_dur;          /* [n_phases] ticks */
  const int32_t* phase_state_off;    /* [n_phases] into state_chars */
  const uint8_t* state_chars;
  /* controlled signals */
  const int32_t* sig_tls;
  const int32_t* sig_n_green;
  const int32_t* sig_yellow_off;     /* [n_signals+1] into yellow_idx (n_green*n_green each) */
  const int32_t* yellow_idx;         /* yellow_dict["i_j"] or -1 (traffic_signal.py:7-24) */
  const int32_t* sig_lane_off;       /* [n_signals+1] */
  const int32_t* sig_lane;
  const int32_t* mv_off;             /* [n_signals*12+1] */
  const int32_t* mv_lane;            /* slot in the signal's lane list */
  const int32_t* mvo_off;            /* [n_signals*12+1] */
  const int32_t* mvo_sig;
  const int32_t* mvo_slot;
  const int32_t* out_off;            /* [n_signals+1] */
  const int32_t* out_sig;
  const int32_t* out_slot;
  /* demand */
  const float* vtype;                /* [n_vtypes][8]: length,minGap,accel,decel,tau,sigma,maxSpeed,speedDev */
  const int32_t* vtype_bit;
  const int32_t* route_off;          /* [n_routes+1] */
  const int32_t* route_edge;
  const int32_t* route_mask;
  const int32_t* origin_lane;        /* [n_origins] */
  const int32_t* origin_off;         /* [n_origins+1] into trip_* (0s when synthetic) */
  const float* trip_depart;          /* seconds after begin, sorted within origin */
  const int32_t* trip_route;
  const int32_t* trip_vtype;
  const int32_t* trip_file;          /* index in the route file (tripinfo order) */
  const int32_t* trip_depart_pos;    /* 0: departPos="base" (back of the origin lane); 1: departPos="random_free" (arterial4x4's
                                      * route files): up to ten random positions on the origin lane are tried for one
                                      * where the vehicle fits between its neighbours, then the base rule */
  /* synthetic Bernoulli-per-tick demand; used when synthetic != 0 */
  const int32_t* origin_rate;        /* P(insert request per tick) * 2^24 */
  const int32_t* origin_route_off;   /* [n_origins+1] into origin_route */
  const int32_t* origin_route;
  /* insertion safety: lanes within 60 m upstream of each origin lane */
  const int32_t* origin_watch_off;   /* [n_origins+1] */
  const int32_t* origin_watch_lane;
  const float* origin_watch_dist;    /* end of watch lane -> start of origin lane (m) */
  const int32_t* origin_watch_owner; /* [n_watch] origin index of each entry */
  /* lane-change safety: lanes within 60 m upstream of each lane of a multi-lane edge (empty otherwise) */
  const int32_t* lane_watch_off;     /* [n_lanes+1] */
  const int32_t* lane_watch_lane;    /* [n_lane_watch] */
  const float* lane_watch_dist;      /* end of watch lane -> start of the lane (m) */
  /* parameters */
  int32_t synthetic;
  int32_t synthetic_vtype;
  int32_t step_length;               /* MultiSignal(step_length) ticks per env step */
  int32_t yellow_length;
  int32_t end_tick;                  /* (end_time - begin) */
  float max_distance;                /* detector range, traffic_signal.py:238-247 */
  float sigma_override;              /* <0: use vType sigma */
  float speed_dev_override;          /* <0: use vType speedDev */
  int32_t vcap;                      /* max concurrently active vehicles per instance */
  int32_t lane_change;               /* 0 disables the lane-change decision */
  int32_t record_trips;              /* keep a per-trip arrival record (tripinfo output); trip-table demand only */
  int32_t tile_vcap;                 /* vehicles of an instance the env-step kernel keeps in shared memory (0: chosen from
                                      * the map's size).  Performance only: an instance that outgrows the tile is stepped
                                      * by the overflow pass out of a global-memory workspace with room for all `vcap`
                                      * vehicles; results do not depend on this value. */
} RsScenario;

/* Optional observation tensors, off by default (rs_select_outputs): each costs its stores every env step. */
#define RS_OUT_DRQ 1            /* states.drq          (states.py:6-31)   */
#define RS_OUT_DRQ_NORM 2       /* states.drq_norm     (states.py:34-59)  */
#define RS_OUT_MPLIGHT_FULL 4   /* states.mplight_full (states.py:83-113) */
/* which tensor rs_env_step_host[_async] copies back as `h_obs` (rs_set_host_obs) */
#define RS_HOSTOBS_MPLIGHT 0       /* [N, S, 13] */
#define RS_HOSTOBS_WAVE 1          /* [N, S, 12] */
#define RS_HOSTOBS_DRQ_NORM 2      /* [N, n_sig_lanes, 5] */
#define RS_HOSTOBS_DRQ 3           /* [N, n_sig_lanes, 5] */
#define RS_HOSTOBS_MPLIGHT_FULL 4  /* [N, S, 49] */

/* Borrowed device pointers, valid until the next mutating call.  [N,...] row-major. */
typedef struct RsObsView {
  int32_t n_env, n_signals, n_sig_lanes;
  const float* lane_queue;       /* [N, n_sig_lanes]  Signal.observe 'queue' */
  const float* lane_approach;    /* 'approach' */
  const float* lane_total_wait;  /* 'total_wait' */
  const float* lane_max_wait;    /* 'max_wait' */
  const float* lane_speed_sum;   /* sum of vehicle speeds (states.drq*) */
  const int32_t* phase;          /* [N, n_signals]  Signal.phase */
  const float* mplight;          /* [N, n_signals, 13]  states.mplight */
  const float* wave;             /* [N, n_signals, 12]  states.wave */
  const float* reward_wait;      /* [N, n_signals]  rewards.wait */
  const float* reward_wait_norm; /* rewards.wait_norm */
  const float* reward_pressure;  /* rewards.pressure */
  const int32_t* sig_queue_len;  /* [N, n_signals] calc_metrics queue_lengths */
  const int32_t* sig_max_queue;  /* calc_metrics max_queues */
  const float* lane_arrivals;    /* [N, n_sig_lanes]  detected vehicles the signal had NOT seen at its previous observe:
                                  * per lane |vehicles ∩ full_observation['arrivals']| (traffic_signal.py:217-224); with
                                  * queue + approach this gives len(arrivals) / len(departures) (rewards.py:96-106) */
  /* optional (RS_OUT_* bits of rs_select_outputs; stale / zero when not selected).  Rows are the signal's lanes in the
   * order of Signal.lanes, signals back to back: a signal's [1, n_lanes, 5] block is rows sig_lane_off[s] .. of this. */
  const float* drq;              /* [N, n_sig_lanes, 5] = [lane index == phase, approach, total_wait, queue, speed sum] */
  const float* drq_norm;         /* [N, n_sig_lanes, 5] = [same one-hot, approach/28, total_wait/28, queue/28, speed sum/20/28] */
  const float* mplight_full;     /* [N, n_signals, 49] = [phase, 12 x (pressure, sum(total_wait/28), speed sum of the movement's
                                  * LAST lane (states.py:97 resets it per lane), sum(approach/28))] */
} RsObsView;

/* Per-instance episode statistics (utils/readXML.py:27-77 inputs). */
typedef struct RsStats {
  int32_t tick;            /* simulation.getTime() - begin */
  int32_t n_active;
  int32_t n_inserted;
  int32_t n_arrived;
  int32_t n_backlog;       /* departed-time reached but not yet inserted */
  int32_t anomalies;       /* ordering violations detected (must stay 0) */
  float sum_delay_arrived; /* sum(timeLoss + departDelay) over finished trips */
  float sum_delay_running; /* same over vehicles still in the net */
  float sum_delay_pending; /* (now - depart) over not-yet-inserted trips */
  float sum_duration_arrived;
  float sum_wait_arrived;  /* sum of tripinfo waitingTime (seconds with speed < 0.1 m/s) over finished trips */
  int32_t sum_active_ticks; /* sum over ticks of n_active (for the roofline V-bar) */
  int32_t n_cap_refused;   /* insertions that had room on the road but were put off because the instance already held
                            * `vcap` vehicles (one count per origin and tick).  SUMO has no such limit: a run in which
                            * this is not 0 was truncated by the capacity setting, not by traffic. */
} RsStats;

typedef struct RsSim RsSim;

const char* rs_last_error(void);
int rs_abi_version(void);

/* traci.start(...) for n_env lock-step instances on CUDA device `device`. */
int rs_create(const RsScenario* sc, int32_t n_env, int32_t device, uint64_t seed, RsSim** out);
/* traci.close() */
int rs_destroy(RsSim* sim);
/* MultiSignal.reset(): fresh episode state for every instance (per-instance RNG key = seed + global id).
 * `first_env_id` is the global id of local instance 0 (multi-GPU sharding keeps results invariant). */
int rs_reset(RsSim* sim, uint64_t seed, int64_t first_env_id, void* stream);
/* The demand of the next episode: per origin lane, the range [h_origin_off[o], h_origin_off[o + 1]) of the trip table
 * (trip_depart / trip_route / trip_vtype, departure order within the range) that may depart -- the route file of the run
 * (`route + '_' + str(self.run) + '.rou.xml'`, multi_signal.py:124) when the trip table holds several files back to
 * back.  Call before rs_reset; all instances share the window.  [n_origins + 1] host ints, ranges inside the table. */
int rs_set_demand_window(RsSim* sim, const int32_t* h_origin_off);
/* trafficlight.setPhase(id, idx) for all instances: phase[N,S] device ptr, mask[N,S] (may be NULL) */
int rs_set_phase(RsSim* sim, const int32_t* d_phase, const uint8_t* d_mask, void* stream);
/* simulationStep() x n_ticks */
int rs_tick(RsSim* sim, int32_t n_ticks, void* stream);
/* Signal.observe + states + rewards into the internal buffers (reset-time observe). */
int rs_observe(RsSim* sim, void* stream);
/* MultiSignal.step fused: prep_phase -> yellow ticks -> set_phase -> green ticks -> observe.
 * d_actions: [N,S] int32 green-phase indices on the device. */
int rs_env_step(RsSim* sim, const int32_t* d_actions, void* stream);
/* Same through HOST buffers (the end-to-end call): copies actions H2D, steps, copies
 * obs/reward back; h_obs [N,S,13] mplight (or the tensor chosen with rs_set_host_obs), h_reward [N,S] (kind: 0 wait,
 * 1 wait_norm, 2 pressure). */
int rs_env_step_host(RsSim* sim, const int32_t* h_actions, float* h_obs, float* h_reward, int32_t reward_kind);
/* Asynchronous form of rs_env_step_host: enqueues actions H2D -> env step -> obs/reward D2H on `stream` and
 * returns at once; rs_wait() blocks until that step's results are in h_obs / h_reward.  The three buffers must
 * stay valid (and h_actions unchanged) until rs_wait returns.  Two sims that each hold half of a batch, driven
 * alternately on two streams, let the host-side agent of one half overlap the device step of the other -- the
 * vectorised-env form of the reference's act -> step loop (main.py:run_trial).  One pending step per sim. */
int rs_env_step_host_async(RsSim* sim, const int32_t* h_actions, float* h_obs, float* h_reward, int32_t reward_kind,
                           void* stream);
int rs_wait(RsSim* sim);
/* Batched MaxPressure / MaxWave action selection on the device (agents/maxwave.py:18-38 over
 * states.mplight[1:] / states.wave): writes [N,S] actions.  pairs [n_pairs,2]; order [S, n_pairs, 2] =
 * (pair index, action) in the reference's evaluation order (iteration order of valid_acts[signal]),
 * pair index -1 terminates a row; ties resolve to the first maximum like the reference's strict '>'. */
int rs_policy_maxpressure(RsSim* sim, const int32_t* h_pairs, int32_t n_pairs, const int32_t* h_order,
                          int32_t use_wave, int32_t* d_actions_out, void* stream);
/* Host-side agent front-end: the same selection over observation rows that are already in HOST memory (what
 * rs_env_step_host / rs_wait returned), for callers that run act -> step from host buffers.  h_obs [n_env,
 * n_signals, obs_dim]; `skip` leading entries of a row are not pressures (1 for states.mplight -- the phase --,
 * 0 for states.wave; agents/maxpressure.py:15-17).  No device work, no RsSim. */
int rs_host_agent_wave(const float* h_obs, int32_t n_env, int32_t n_signals, int32_t obs_dim, int32_t skip,
                       const int32_t* pairs, int32_t n_pairs, const int32_t* order, int32_t* h_actions);
/* ---- agent front ends of shared-policy / exploring agents ----
 * FRAP, the Q-network of MPLight (agents/mplight.py:43-131): parameters as host float arrays in the reference module's
 * own names and PyTorch layouts (a state_dict's tensors, C-contiguous); demand_shape = 1 (agent_config.py MPLight). */
typedef struct RsFrapParams {
  const float* p_weight;                  /* [2, 4]   */
  const float* d_weight;                  /* [4, 1]   */
  const float* d_bias;                    /* [4]      */
  const float* lane_embedding_weight;     /* [16, 8]  */
  const float* lane_embedding_bias;       /* [16]     */
  const float* lane_conv_weight;          /* [20, 32, 1, 1] */
  const float* lane_conv_bias;            /* [20]     */
  const float* relation_embedding_weight; /* [2, 4]   */
  const float* relation_conv_weight;      /* [20, 4, 1, 1] */
  const float* relation_conv_bias;        /* [20]     */
  const float* hidden_layer_weight;       /* [20, 20, 1, 1] */
  const float* hidden_layer_bias;         /* [20]     */
  const float* before_merge_weight;       /* [1, 20, 1, 1] */
  const float* before_merge_bias;         /* [1]      */
} RsFrapParams;
/* Upload the network and the action tables: pairs [n_pairs, 2] (signal_configs[map]['phase_pairs'], n_pairs <= 16),
 * order [S, n_pairs, 2] as for rs_policy_maxpressure (valid_acts in the reference's scan order). */
int rs_frap_load(RsSim* sim, const RsFrapParams* params, const int32_t* h_pairs, int32_t n_pairs, const int32_t* h_order);
/* Q-values + greedy valid action for every (instance, signal) row.  d_obs: [n_env_rows, S, 13] states.mplight rows on
 * the device, or NULL for the sim's own observation (then n_env_rows is ignored) -- a rank that evaluates the shared
 * policy for all ranks passes the all-gathered tensor.  d_actions_out [n_env_rows, S] (NULL: the sim's own action
 * buffer, consumed by rs_env_step_policy); d_q_out [n_env_rows, S, n_pairs] or NULL. */
int rs_policy_frap(RsSim* sim, const float* d_obs, int32_t n_env_rows, int32_t* d_actions_out, float* d_q_out, void* stream);
/* Uniform random green phase per signal (epsilon = 1 exploration, agents/pfrl_dqn.py:57-70): Philox keyed by (seed, global
 * instance id, signal, instance tick), so the draw does not depend on the batch or the GPU the instance runs on. */
int rs_policy_random(RsSim* sim, uint64_t seed, int32_t* d_actions_out, void* stream);
/* One whole agent step as a single CUDA-graph launch: the policy kernel over the sim's current observation
 * (RS_POLICY_*; tables from rs_policy_maxpressure / rs_frap_load must have been uploaded by one ordinary call
 * before) followed by the fused env step on its actions.  The graph is captured on first use and re-captured when the
 * episode seed or the output selection changes. */
#define RS_POLICY_MAXPRESSURE 1
#define RS_POLICY_MAXWAVE 2
#define RS_POLICY_FRAP 3
#define RS_POLICY_RANDOM 4
int rs_env_step_policy(RsSim* sim, int32_t policy, uint64_t policy_seed, void* stream);
int rs_get_obs(RsSim* sim, RsObsView* out);
int rs_get_stats(RsSim* sim, RsStats* h_out /* [N] */);
/* Dump one instance's vehicles to host arrays (TraCI getters of the N=1 facade; parity tests).
 * Arrays have room for vcap entries; returns count through n_out.  lane[i] is the lane index. */
int rs_dump_vehicles(RsSim* sim, int32_t env, int32_t* n_out, int32_t* lane, float* pos, float* speed,
                     float* accel, float* wait, float* rwait, float* tloss, int32_t* vid, int32_t* vtype,
                     int32_t* route, int32_t* cursor, float* sf, int32_t* depart, float* acc_wait);
int rs_get_phases(RsSim* sim, int32_t env, int32_t* h_tls_phase /* [n_tls] */);
/* `--tripinfo-output` (multi_signal.py:127-129): per-trip arrival records of one instance, indexed by trip
 * (order of trip_depart/trip_route).  arrival[i] < 0: not arrived.  Needs RsScenario.record_trips. */
int rs_get_trip_records(RsSim* sim, int32_t env, int32_t* h_arrival_tick, int32_t* h_depart_tick,
                        float* h_time_loss, int32_t* h_depart_delay, float* h_waiting_time);
/* Select the optional observation tensors (RS_OUT_* bits) the observe step writes from now on. */
int rs_select_outputs(RsSim* sim, int32_t mask);
/* Select the tensor rs_env_step_host[_async] returns in h_obs (RS_HOSTOBS_*; default mplight); the matching RS_OUT_*
 * bit is switched on.  floats_per_instance (may be NULL) receives the row size of h_obs. */
int rs_set_host_obs(RsSim* sim, int32_t kind, int32_t* floats_per_instance);
int64_t rs_kernel_launches(RsSim* sim);
/* launch shape chosen for this scenario (diagnostics / bench `config`): threads per instance, instances per CTA,
 * CTAs in the persistent grid, dynamic shared memory per CTA, tile buffers in shared memory (2 = ping-pong,
 * 1 = re-sort through registers, 0 = the vehicle store does not fit shared memory and lives in a per-CTA global-memory
 * workspace).  Any pointer may be NULL. */
int rs_get_launch_shape(RsSim* sim, int32_t* threads_per_instance, int32_t* instances_per_cta, int32_t* grid_ctas,
                        int32_t* smem_bytes_per_cta, int32_t* tile_buffers);
/* the vehicle store and the tile (diagnostics / bench `config`): vehicles per instance in the fast pass's tile, in the
 * HBM store, and in the tile a CTA lays over its whole shared memory to step an instance that outgrew its slot again
 * (0: none, one instance per CTA); whether an overflow pass out of the global-memory workspace exists for what outgrows
 * those; how many instances the LAST launch stepped with a whole CTA on that tile (from the heavy list, or again after
 * outgrowing their slot) / deferred to the overflow pass (these two synchronise the device).  Any pointer may be NULL. */
int rs_get_tile_info(RsSim* sim, int32_t* tile_vcap, int32_t* store_vcap, int32_t* redo_vcap, int32_t* has_overflow_pass,
                     int32_t* last_redone, int32_t* last_deferred);
/* device time (ms) of the last rs_env_step's kernels, CUDA events on the launching stream */
int rs_last_step_ms(RsSim* sim, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* RESCO_B200_H */
