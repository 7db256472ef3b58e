/* wf_oracle.h -- CPU restatement of the reference fire-spread step.
 *
 * TEST INFRASTRUCTURE.  Nothing in the product (wildfire_control_python_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity pin: validated step by step against the unmodified Python reference
 * (oracle/validate_oracle.py, oracle/gen_golden.py -> tests/golden/).
 */
#ifndef WF_ORACLE_H
#define WF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors Simulation/constants.py:30-47 (METADATA) and Simulation/utility.py:94-102 (grass). */
typedef struct wfo_config {
    int32_t width, height;
    int32_t n_actions;
    int32_t a_speed;
    int32_t allow_dig_toggle;
    int32_t make_rivers;
    int32_t containment_wins;
    int32_t wind_random;      /* METADATA['wind'] == "random" */
    int32_t wind_x, wind_y;   /* else METADATA['wind'] = [wind_speed, (wind_x, wind_y)] */
    int32_t fuel;             /* grass['fuel'] */
    int32_t radius;           /* grass['radius'] */
    int32_t extra_ignitions;  /* World.set_fire_to() calls right after reset() */
    int32_t _pad;
    double wind_speed;
    double death_penalty, contained_bonus, default_reward;
    double heat, threshold;   /* grass['heat'], grass['threshold'] */
    uint64_t seed;
} wfo_config;

typedef struct wfo_env wfo_env;

void wfo_default_config(wfo_config* cfg, int size);
void wfo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

wfo_env* wfo_create(const wfo_config* cfg, int64_t env_id);
void wfo_destroy(wfo_env* e);
/* World.reset(); episode counter += 1; start/wind/river from the Philox RESET stream. */
void wfo_reset(wfo_env* e);
/* Same, but start cell forced (no stream draws for it) -- for scripted tests. */
void wfo_reset_at(wfo_env* e, int ax, int ay);
/* ForestFire.step(action).  obs_u8 may be NULL; else W*H*3 bytes, index (x*H+y)*3+c.
 * Returns 0, or -1 when the reference would raise (move with no agent, Q8). */
int wfo_step(wfo_env* e, int action, uint8_t* obs_u8, double* reward, int* done);
/* The action the shared ACTION stream prescribes for the env's current (episode, t). */
int wfo_stream_action(const wfo_env* e);
/* The reference's heuristic walk policy (DQN.choose_randomwalk_action, DQN.py:353-389), POLICY stream. */
int wfo_policy_action(const wfo_env* e);
void wfo_get_obs(const wfo_env* e, uint8_t* obs_u8);

/* Canonical planes, each W*H, index x*H+y.  Any pointer may be NULL. */
void wfo_get_planes(const wfo_env* e, uint8_t* type, uint8_t* burning, uint8_t* fm_inf,
                    int32_t* fuel, double* temp, uint8_t* apos);
/* scalars: [0]=alive [1]=ax [2]=ay [3]=dead [4]=digging [5]=running [6]=fire_at_border
 * [7]=n_border_points [8]=episode [9]=t [10]=a_speed_iter [11]=wind_x [12]=wind_y [13]=n_burning */
void wfo_get_scalars(const wfo_env* e, int32_t out[16]);
double wfo_get_wind_speed(const wfo_env* e);
void wfo_set_a_speed_iter(wfo_env* e, int v); /* the reference's counter is process-global (Q8) */
/* Directional heat quanta: d = 0 N(0,-1) 1 S(0,+1) 2 E(+1,0) 3 W(-1,0) (displacement source->target). */
void wfo_get_coef(const wfo_env* e, double coef[4]);
/* World.set_fire_to((x, y)) -- environment.py:233-246. */
void wfo_set_fire_to(wfo_env* e, int x, int y);
/* Overwrite planes (river parity: upload the reference's reset map). NULL = keep. */
void wfo_set_planes(wfo_env* e, const uint8_t* type, const uint8_t* burning, const uint8_t* fm_inf,
                    const int32_t* fuel, const double* temp);
void wfo_set_agent(wfo_env* e, int alive, int ax, int ay, int visible, int dead, int digging);

/* Throughput driver (bench.py --impl reference / cpu_baseline): n_envs independent envs,
 * n_steps each, actions from the ACTION stream, reset on done; n_threads pthreads over envs.
 * Returns total env-steps executed; *checksum accumulates rewards (keeps the work alive). */
int64_t wfo_rollout(const wfo_config* cfg, int64_t env_id_base, int n_envs, int n_steps,
                    int n_threads, double* checksum, int64_t* episodes);

/* Persistent batch for bench.py --impl reference: envs live across calls. */
typedef struct wfo_batch wfo_batch;
wfo_batch* wfo_batch_create(const wfo_config* cfg, int64_t env_id_base, int n_envs, int n_threads);
void wfo_batch_destroy(wfo_batch* b);
int64_t wfo_batch_step(wfo_batch* b, int n_steps, double* checksum, int64_t* episodes);

#ifdef __cplusplus
}
#endif
#endif
