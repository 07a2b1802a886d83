/* TEST INFRASTRUCTURE ONLY — the oracle.  Never imported, linked or executed by the product
 * (interactive_highway_simulation_b200/, libmcsim_b200.so); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker/baseline.
 *
 * mcsim_oracle: scalar C++ restatement of the reference's inner Monte-Carlo simulator
 * (mc_sim_highway.so: Simulation::run 0x10e790 and everything it calls).  The reference ships
 * no source for that binary; this file restates its disassembly function by function (each
 * function below cites the address range it follows).  PARITY PINNED: tests/test_oracle_vs_ref.py
 * drives this restatement and the reference's own machine code (oracle/_ref/ref_harness) with the
 * same scenes and the same injected random stream and requires bit-identical per-step states and
 * rollout statistics; the fixtures under tests/golden/ were produced by the reference binary.
 */
#ifndef MCSIM_ORACLE_H
#define MCSIM_ORACLE_H
#include "../include/mcsim.h"

#ifdef __cplusplus
extern "C" {
#endif

/* math_mode: 0 = include/mcsim_math.h (mcs_exp / mcs_pow4), 1 = platform libm (exp / pow) */
void orc_set_math_mode(int mode);

/* One fresh Simulation + run() per rollout r in [first, first+count): stream key =
 * mcs_rollout_key(seed, r); draws 0..N-1 are the Agent-ctor aggressive flags in input order.
 * status_out[count]; states_out (may be NULL) = [count][n_agents][steps+1][4], n_states_out[count]. */
int orc_rollouts(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                 const mcs_candidate* cand, const mcs_sim_param* param, uint64_t seed, int64_t first, int64_t count,
                 mcs_status* status_out, double* states_out, int32_t* n_states_out);

/* same, rollouts spread over n_threads host threads (independent streams → order-free) */
int orc_rollouts_mt(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                    const mcs_candidate* cand, const mcs_sim_param* param, uint64_t seed, int64_t first, int64_t count,
                    int n_threads, mcs_status* status_out);

/* runMTimes/runMultiThreads two-level mean over groups × n_per_group statuses (fixed order) */
int orc_reduce_features(const mcs_status* statuses, int groups, int n_per_group, int steps, mcs_feature* out);

/* agent visit order / corridor order / neighbour ids the reference derives (hash-table iteration order) */
int orc_scene_orders(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                     int32_t* agent_order, int32_t* corridor_order, int32_t* left_ids, int32_t* right_ids);

/* Python-facing single decisions on a freshly built Environment (draw stream key = seed, counter from 0,
 * draws 0..N-1 again the ctor flags) */
int orc_decide_mobil(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                     int32_t agent_id, const mcs_action* idm_action, int require_rss, uint64_t key,
                     mcs_action_with_type* out);
int orc_decide_fixed_direction(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                               int32_t agent_id, int32_t action, uint64_t key, mcs_action* out);
int orc_decide_fixed_gap(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                         int32_t agent_id, int32_t direction, int32_t gap_id, uint64_t key, mcs_action* out);
int orc_decide_closest_gap(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                           int32_t agent_id, int32_t direction, uint64_t key, mcs_action* out);

/* timeToReachGap(xBack, vBack, xFront, vFront, xEgo, vEgo, vTarget), p = the 7 arguments */
double orc_time_to_reach_gap(const double* p);

#ifdef __cplusplus
}
#endif
#endif
