/*
 * vine_oracle.h — CPU ORACLE for the Vine5LinkMovingBase hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (vine_robot_isaacgymenvs_b200/) never does.
 *
 * Parity status:
 *   - task logic (action path V5:922-945, 984-1005; actuation V5:1028-1106; observations
 *     V5:1339-1390; reward V5:1218-1248,1470-1537; resets V5:1540-1558, 774-914; step
 *     sequencing VT:319-380): PINNED against the reference's own Python functions, executed
 *     from /root/reference by tests/golden/generate_golden.py (fixtures in tests/golden/).
 *   - dynamics + contact (gym.simulate, VT:356): the reference delegates to the closed
 *     Isaac Gym / PhysX binary, which is not available and for which the reference holds no
 *     golden vector: PARITY UNPINNED.  The oracle restates the URDF model.
 */
#ifndef VINE_ORACLE_H_
#define VINE_ORACLE_H_

#include <stdint.h>
#include "../include/vine_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Arrays follow the reference's torch tensors (row-major AoS). */
typedef struct OracleArrays {
  int64_t n;
  int64_t global_env_offset;
  uint64_t seed;
  /* persistent state */
  float* dof_pos;              /* [N,6] */
  float* dof_vel;              /* [N,6] */
  float* tip_body;             /* [N,3] rigid-body tip position as of the last simulate */
  float* tipvel_body;          /* [N,3] */
  float* cart_body_y;          /* [N] */
  float* cart_body_vy;         /* [N] */
  float* target;               /* [N,3] */
  float* object_info;          /* [N,2] */
  float* smoothed;             /* [N] */
  float* prev_cart_vel;        /* [N] */
  float* prev_cart_vel_error;  /* [N] */
  float* lip_force;            /* [N] */
  float* history;              /* [N,D,2] oldest first */
  float* agg_rew;              /* [N] */
  int64_t* step_count;         /* [N] */
  /* VecTask io */
  const float* actions;        /* [N,2] */
  float* obs;                  /* [N,O] */
  float* obs_clamped;          /* [N,O] or NULL */
  float* rew;                  /* [N] */
  int64_t* reset;              /* [N] */
  int64_t* progress;           /* [N] */
  uint8_t* timeout;            /* [N] */
  /* outputs of the last step (may be NULL) */
  float* u_rail;
  float* u_fpam;
  float* prev_u_rail;
  float* rail_force;
  float* reward_matrix;        /* [N,13] */
} OracleArrays;

int oracle_num_observations(int observation_type);
void oracle_config_defaults(VineConfig* cfg);

/* Philox4x32-10: ctr=(gid, site, step, block), key=seed -> out[4]. */
void oracle_philox(uint64_t seed, uint32_t gid, uint32_t site, uint32_t step, uint32_t block,
                   uint32_t out[4]);
/* Same uniform / normal conversions the product uses; for fixture generation. */
void oracle_dynamics_uniforms(uint64_t seed, uint32_t gid, uint32_t step, uint32_t sim_i, float out[24]);
void oracle_uniform4(uint64_t seed, uint32_t gid, uint32_t site, uint32_t step, uint32_t block,
                     float out[4]);
void oracle_normal4(uint64_t seed, uint32_t gid, uint32_t site, uint32_t step, uint32_t block,
                    float out[4]);

/* Whole control step (VT:319-380) for all envs; use_f64 selects the dynamics precision.
 * nthreads<=0: all OpenMP threads. */
int oracle_step(const VineConfig* cfg, OracleArrays* a, int use_f64, int nthreads);
int oracle_init(const VineConfig* cfg, OracleArrays* a);
int oracle_reset_idx(const VineConfig* cfg, OracleArrays* a, const int64_t* env_ids, int64_t n);

/* Function-level restatements (same argument structs as the product ABI, HOST pointers). */
int oracle_pre_physics(const VineConfig* cfg, int64_t n, const VinePrePhysicsIO* io);
int oracle_actuation(const VineConfig* cfg, int64_t n, const VineActuationIO* io);
int oracle_simulate(const VineConfig* cfg, int64_t n, const VineSimulateIO* io, int use_f64);
int oracle_post_physics(const VineConfig* cfg, int64_t n, const VinePostPhysicsIO* io);
int oracle_gae(const float* rewards, const float* values, const float* dones,
               const float* last_values, const float* last_dones, int64_t horizon,
               int64_t num_envs, double gamma, double tau, float* advantages, float* returns);

/* Diagnostics for the dynamics model (f64): joint-space mass matrix in the reference's
 * relative coordinates, and total mechanical energy. q,qd: [6]. */
void oracle_mass_matrix(const double q[6], double M[36]);
double oracle_energy(const VineConfig* cfg, const double q[6], const double qd[6]);

#ifdef __cplusplus
}
#endif
#endif
