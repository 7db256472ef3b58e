/* wf_oracle.c -- plain-C restatement of the reference's environment step.
 *
 * TEST INFRASTRUCTURE (see wf_oracle.h).  Sequential, one env at a time, float64
 * temperatures, explicit burning list and border-point stack -- i.e. the same
 * data structures and control flow as the Python reference, so that it can be
 * checked line by line against it.  Citations are relative to /root/reference.
 *
 * Known, documented differences from the Python object (none observable on the
 * bit-exact fields in any validated rollout):
 *   - Python iterates `set`s in hash order; we iterate the burning list in
 *     insertion order.  Only the temperature of a cell in the very tick it
 *     ignites depends on that order, and that value is never read again.
 *   - `burning.pop()` in get_reward picks an arbitrary cell; we take the last.
 */
#include "wf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

enum { T_GRASS = 0, T_FIRE = 1, T_BURNT = 2, T_DIRT = 3, T_WATER = 4 }; /* utility.py:128-140 */

struct wfo_env {
    wfo_config cfg;
    int W, H;
    int64_t env_id;
    uint32_t episode; /* incremented by every reset; first reset -> 0 */
    uint32_t t;       /* steps taken in this episode */
    uint32_t draw_k;  /* sequential index into the RESET stream */
    /* env[x, y, layer] of environment.py:38-50, only the layers that ever change */
    uint8_t* type;
    double* temp;
    int32_t* fuel;
    uint8_t* fm_inf; /* fire_mobility == inf */
    uint8_t* apos;   /* agent_pos */
    /* World.burning_cells (set) */
    int32_t* burn_list;
    int32_t n_burn;
    uint8_t* in_burn;
    /* World.border_points (deque used as a stack: pop()/append(), environment.py:355,377) */
    int32_t* bp;
    int32_t n_bp;
    /* World.agents[0] (single agent) */
    int alive, ax, ay, dead, digging;
    double wind_speed;
    int wind_x, wind_y;
    int running, fire_at_border;
    int a_speed_iter; /* METADATA['a_speed_iter'] -- NOT reset by reset() (Q8, forest_fire.py:40-43) */
    /* scratch */
    uint8_t* comp;
    int32_t* queue;
};

/* ---------------- Philox4x32-10 (Salmon et al., Random123) ---------------- */
void wfo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void stream_block(const wfo_env* e, uint32_t index, uint32_t stream, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)e->env_id, e->episode, index, stream};
    uint32_t key[2] = {(uint32_t)(e->cfg.seed & 0xffffffffu), (uint32_t)(e->cfg.seed >> 32)};
    wfo_philox4x32_10(ctr, key, out);
}

/* next draw of the RESET stream (stream 0) -- stands in for np.random.choice / random.randint */
static uint32_t reset_draw(wfo_env* e) {
    uint32_t w[4];
    stream_block(e, e->draw_k >> 2, 0u, w);
    return w[(e->draw_k++) & 3u];
}

void wfo_default_config(wfo_config* c, int size) { /* constants.py:30-47, utility.py:94-102 */
    memset(c, 0, sizeof(*c));
    c->width = c->height = size;
    c->n_actions = 4;
    c->a_speed = 1;
    c->wind_speed = 0.54;
    c->wind_x = c->wind_y = 0;
    c->death_penalty = -1000.0;
    c->contained_bonus = 1000.0;
    c->default_reward = -1.0;
    c->heat = 0.3;
    c->threshold = 3.0;
    c->fuel = 20;
    c->radius = 1;
}

#define IDX(e, x, y) ((x) * (e)->H + (y)) /* env[x, y]: x is the slow axis */

wfo_env* wfo_create(const wfo_config* cfg, int64_t env_id) {
    wfo_env* e = (wfo_env*)calloc(1, sizeof(wfo_env));
    e->cfg = *cfg;
    e->W = cfg->width;
    e->H = cfg->height;
    e->env_id = env_id;
    int n = e->W * e->H;
    e->type = (uint8_t*)calloc(n, 1);
    e->temp = (double*)calloc(n, sizeof(double));
    e->fuel = (int32_t*)calloc(n, sizeof(int32_t));
    e->fm_inf = (uint8_t*)calloc(n, 1);
    e->apos = (uint8_t*)calloc(n, 1);
    e->burn_list = (int32_t*)calloc(n, sizeof(int32_t));
    e->in_burn = (uint8_t*)calloc(n, 1);
    e->bp = (int32_t*)calloc(2 * (size_t)(2 * e->W + 2 * e->H) + 8, sizeof(int32_t));
    e->comp = (uint8_t*)calloc(n, 1);
    e->queue = (int32_t*)calloc(n, sizeof(int32_t));
    e->episode = 0xFFFFFFFFu; /* first reset() wraps to episode 0 */
    e->a_speed_iter = cfg->a_speed;
    return e;
}

void wfo_destroy(wfo_env* e) {
    if (!e) return;
    free(e->type); free(e->temp); free(e->fuel); free(e->fm_inf); free(e->apos);
    free(e->burn_list); free(e->in_burn); free(e->bp); free(e->comp); free(e->queue);
    free(e);
}

/* World.inbounds -- environment.py:225-226 */
static int inbounds(const wfo_env* e, int x, int y) { return 0 <= x && x < e->W && 0 <= y && y < e->H; }
/* World.is_burning -- environment.py:249-251 (a TYPE test, not set membership: Q7) */
static int is_burning(const wfo_env* e, int x, int y) { return e->type[IDX(e, x, y)] == T_FIRE; }
/* World.is_burnable -- environment.py:254-257 */
static int is_burnable(const wfo_env* e, int x, int y) { return e->type[IDX(e, x, y)] == T_GRASS; }

/* World.set_fire_to -- environment.py:233-246 */
void wfo_set_fire_to(wfo_env* e, int x, int y) {
    int i = IDX(e, x, y);
    if (e->temp[i] < e->cfg.threshold) e->temp[i] = e->cfg.threshold + 1;
    e->type[i] = T_FIRE;
    if (!e->in_burn[i]) { /* set.add */
        e->in_burn[i] = 1;
        e->burn_list[e->n_burn++] = i;
    }
    if (x == 0 || x == e->W - 1 || y == 0 || y == e->H - 1) e->fire_at_border = 1;
}

/* World.reset_border_points -- environment.py:215-222 (the [HEIGHT-1, y] is literal) */
static void reset_border_points(wfo_env* e) {
    e->n_bp = 0;
    for (int x = 0; x < e->W; ++x) {
        e->bp[2 * e->n_bp] = x; e->bp[2 * e->n_bp + 1] = 0; e->n_bp++;
        e->bp[2 * e->n_bp] = x; e->bp[2 * e->n_bp + 1] = e->H - 1; e->n_bp++;
    }
    for (int y = 0; y < e->H; ++y) {
        e->bp[2 * e->n_bp] = 0; e->bp[2 * e->n_bp + 1] = y; e->n_bp++;
        e->bp[2 * e->n_bp] = e->H - 1; e->bp[2 * e->n_bp + 1] = y; e->n_bp++;
    }
}

/* Agent.dig -- environment.py:123-133 */
static void agent_dig(wfo_env* e) {
    if (!e->digging) return;
    int i = IDX(e, e->ax, e->ay);
    if (e->type[i] != T_DIRT) {
        e->type[i] = T_DIRT;
        e->fm_inf[i] = 1;
    }
}

/* reset_map -- environment.py:59-95 */
static void reset_map(wfo_env* e) {
    int n = e->W * e->H;
    for (int i = 0; i < n; ++i) {
        e->temp[i] = 0.0;
        e->fuel[i] = e->cfg.fuel;
        e->fm_inf[i] = 0;
        e->apos[i] = 0;
        e->type[i] = T_GRASS;
    }
    if (!e->cfg.make_rivers) return;
    const int W = e->W, H = e->H;
    const int fx = W / 2, fy = H / 2;                  /* :73, utility.py:61-64 */
    int river_x = (int)(reset_draw(e) % (uint32_t)W); /* :75 */
    int river_y = 1 + (int)(reset_draw(e) % 3u);      /* :77  choice([1,2,3]) */
    while (river_y < H - (1 + (int)(reset_draw(e) % 3u))) { /* :79 */
        int i = IDX(e, river_x, river_y);
        e->type[i] = T_WATER; /* :81 */
        e->fm_inf[i] = 1;     /* :85 */
        int new_y = river_y + 1;
        int new_x = river_x + ((reset_draw(e) % 2u) ? -1 : 1); /* :90 choice([1,-1]) */
        for (;;) { /* :91-93 -- Python's chained comparison short-circuits its 2nd draw */
            int lo = 1 + (int)(reset_draw(e) % 3u);
            int chain = 0;
            if (lo <= new_x) {
                int hi = W - (1 + (int)(reset_draw(e) % 3u));
                chain = new_x < hi;
            }
            if (chain) break;
            if (new_x == fx && new_y == fy) break;
            new_x = river_x + ((reset_draw(e) % 2u) ? -1 : 1);
        }
        river_x = new_x;
        river_y = new_y;
    }
}

/* circle_points(midx, midy, r) -- utility.py:8-52; returns the number of points */
static int circle_points(int midx, int midy, int r, int (*out)[2]) {
    int n = 0, x = r, y = 0;
    out[n][0] = x + midx; out[n][1] = y + midy; n++;
    if (r > 0) {
        out[n][0] = -x + midx; out[n][1] = -y + midy; n++;
        out[n][0] = y + midx; out[n][1] = -x + midy; n++;
        out[n][0] = -y + midx; out[n][1] = x + midy; n++;
    }
    int P = 1 - r;
    while (x > y) {
        y += 1;
        if (P <= 0) P = P + 2 * y + 1;
        else { x -= 1; P = P + 2 * y - 2 * x + 1; }
        if (x < y) break;
        out[n][0] = x + midx; out[n][1] = y + midy; n++;
        out[n][0] = -x + midx; out[n][1] = y + midy; n++;
        out[n][0] = x + midx; out[n][1] = -y + midy; n++;
        out[n][0] = -x + midx; out[n][1] = -y + midy; n++;
        if (x != y) {
            out[n][0] = y + midx; out[n][1] = x + midy; n++;
            out[n][0] = -y + midx; out[n][1] = x + midy; n++;
            out[n][0] = y + midx; out[n][1] = -x + midy; n++;
            out[n][0] = -y + midx; out[n][1] = -x + midy; n++;
        }
    }
    return n;
}

static void world_reset(wfo_env* e, int force_start, int sx, int sy) {
    e->episode += 1u;
    e->t = 0;
    e->draw_k = 0;
    /* wind -- environment.py:188-193 */
    if (e->cfg.wind_random) {
        static const double speeds[3] = {0.0, 0.7, 0.85};
        e->wind_speed = speeds[reset_draw(e) % 3u];
        e->wind_x = -1 + (int)(reset_draw(e) % 3u);
        e->wind_y = -1 + (int)(reset_draw(e) % 3u);
    } else {
        e->wind_speed = e->cfg.wind_speed;
        e->wind_x = e->cfg.wind_x;
        e->wind_y = e->cfg.wind_y;
    }
    e->running = 1;   /* :196 */
    reset_map(e);     /* :199 */
    e->n_burn = 0;    /* :202 */
    memset(e->in_burn, 0, (size_t)e->W * e->H);
    wfo_set_fire_to(e, e->W / 2, e->H / 2); /* :203, utility.py:61-64 */
    /* agent -- :206-208, utility.py:66-78, Agent.__init__ environment.py:100-113 */
    int ax, ay;
    if (force_start) { ax = sx; ay = sy; }
    else {
        int pts[64][2];
        int radius = 1 + (int)(reset_draw(e) % 3u);          /* utility.py:70 */
        int n = circle_points(e->W / 2, e->H / 2, radius, pts);
        int idx = (int)(reset_draw(e) % (uint32_t)n);        /* utility.py:75 */
        ax = pts[idx][0]; ay = pts[idx][1];
    }
    e->alive = 1; e->ax = ax; e->ay = ay;
    e->apos[IDX(e, ax, ay)] = 1;
    e->dead = 0; e->digging = 1;
    agent_dig(e);
    reset_border_points(e); /* :211 */
    e->fire_at_border = 0;  /* :212 */
    /* extra ignitions through the public World.set_fire_to, after reset() (IGNITE stream) */
    for (int k = 0; k < e->cfg.extra_ignitions; ++k) {
        uint32_t w[4];
        stream_block(e, (uint32_t)k, 2u, w);
        wfo_set_fire_to(e, (int)(w[0] % (uint32_t)e->W), (int)(w[1] % (uint32_t)e->H));
    }
}

void wfo_reset(wfo_env* e) { world_reset(e, 0, 0, 0); }
void wfo_reset_at(wfo_env* e, int ax, int ay) { world_reset(e, 1, ax, ay); }

/* Agent.move -- environment.py:141-155, _direction_to_coords :163-171 */
static void agent_move(wfo_env* e, int direction) {
    static const int DX[4] = {0, 0, 1, -1}, DY[4] = {-1, 1, 0, 0};
    e->apos[IDX(e, e->ax, e->ay)] = 0; /* cleared BEFORE the validity test: Q1 */
    int nx = e->ax + DX[direction], ny = e->ay + DY[direction];
    if (inbounds(e, nx, ny) && e->type[IDX(e, nx, ny)] != T_WATER) { /* traversable :229-230 */
        e->ax = nx; e->ay = ny;
        e->apos[IDX(e, nx, ny)] = 1;
        if (e->digging && !is_burning(e, nx, ny)) agent_dig(e);
        if (is_burning(e, nx, ny)) e->dead = 1;
    }
}

/* World.reduce_fuel -- environment.py:297-307 */
static int reduce_fuel(wfo_env* e, int pos_in_list) {
    int i = e->burn_list[pos_in_list];
    e->fuel[i] -= 1;
    if (e->fuel[i] <= 0) {
        e->type[i] = T_BURNT;
        e->in_burn[i] = 0; /* removed from burn_list by the caller's compaction */
        return 0;
    }
    return 1;
}

/* World.apply_heat_from_to -- environment.py:278-294 with get_distance_and_angle :260-275 */
static void apply_heat_from_to(wfo_env* e, int x, int y, int ox, int oy) {
    int cx = ox - x, cy = oy - y;
    int wx = e->wind_x, wy = e->wind_y;
    double distance = (double)(abs(x - ox) + abs(y - oy));
    double angle = fabs(atan2((double)(wx * cy - wy * cx), (double)(wx * cx + wy * cy)));
    double env_factor = pow(angle + distance, -1.0);
    double calculated_heat = e->wind_speed * e->cfg.heat * env_factor;
    int j = IDX(e, ox, oy);
    e->temp[j] += calculated_heat;
    if (e->temp[j] > e->cfg.threshold) wfo_set_fire_to(e, ox, oy);
}

/* ForestFire.update -- forest_fire.py:85-106 */
static void update(wfo_env* e) {
    /* :87 Agent.is_dead environment.py:116-120 */
    if (e->alive && (e->dead || is_burning(e, e->ax, e->ay))) {
        e->apos[IDX(e, e->ax, e->ay)] = 0;
        e->alive = 0;
    }
    int n0 = e->n_burn; /* :90 iterate over a copy: cells ignited this tick do not burn yet */
    const int R = e->cfg.radius;
    for (int p = 0; p < n0; ++p) {
        int i = e->burn_list[p];
        if (reduce_fuel(e, p)) { /* :95 */
            int cx = i / e->H, cy = i % e->H;
            /* World.get_neighbours -- environment.py:311-326 (Manhattan diamond) */
            for (int dx = -R; dx <= R; ++dx) {
                int rem = R - abs(dx);
                for (int dy = -rem; dy <= rem; ++dy) {
                    if (dx == 0 && dy == 0) continue;
                    int nx = cx + dx, ny = cy + dy;
                    if (inbounds(e, nx, ny) && is_burnable(e, nx, ny)) /* :324 and :99 */
                        apply_heat_from_to(e, cx, cy, nx, ny);
                }
            }
        }
    }
    /* compact the list: drop burnt-out cells (set.remove, environment.py:305) */
    int m = 0;
    for (int p = 0; p < e->n_burn; ++p) {
        int i = e->burn_list[p];
        if (e->in_burn[i]) e->burn_list[m++] = i;
    }
    e->n_burn = m;
    if (!e->alive || e->n_burn == 0) e->running = 0; /* :105-106 */
}

/* Reachability restatement of pyastar.astar_path(grid, start, goal, allow_diagonal=False)
 * (pyastar/astar.cpp:41-116): a cell is entered only if its weight is finite
 * (new_cost = cost + inf is never < inf, :89-90); the START cell's own weight is
 * never read; so "a path exists" <=> goal is in the 4-connected flood from start
 * over finite cells.  start == goal yields an empty path (pyastar.py:52-64). */
static void flood_from(wfo_env* e, int start) {
    int n = e->W * e->H, head = 0, tail = 0;
    memset(e->comp, 0, (size_t)n);
    e->comp[start] = 1;
    e->queue[tail++] = start;
    while (head < tail) {
        int i = e->queue[head++];
        int x = i / e->H, y = i % e->H;
        static const int DX[4] = {-1, 1, 0, 0}, DY[4] = {0, 0, -1, 1};
        for (int k = 0; k < 4; ++k) {
            int nx = x + DX[k], ny = y + DY[k];
            if (!inbounds(e, nx, ny)) continue;
            int j = IDX(e, nx, ny);
            if (e->comp[j] || e->fm_inf[j]) continue;
            e->comp[j] = 1;
            e->queue[tail++] = j;
        }
    }
}
static int path_exists(const wfo_env* e, int start, int gx, int gy) {
    if (!inbounds(e, gx, gy)) return 0; /* the reference would raise (non-square maps) */
    int g = IDX(e, gx, gy);
    return g != start && e->comp[g];
}

/* World.get_reward -- environment.py:342-390 */
static double get_reward(wfo_env* e) {
    if (!e->fire_at_border && e->n_bp && e->n_burn) {
        int nb = e->n_burn; /* burning = set(self.burning_cells) */
        int b = e->burn_list[--nb]; /* burning.pop() */
        flood_from(e, b);
        int ex = e->bp[2 * (e->n_bp - 1)], ey = e->bp[2 * (e->n_bp - 1) + 1]; /* border_points.pop() */
        e->n_bp--;
        while (!path_exists(e, b, ex, ey)) {
            if (e->n_bp == 0) {
                if (nb == 0) return e->cfg.contained_bonus; /* :363-367 (containment_wins is a no-op) */
                reset_border_points(e);
                b = e->burn_list[--nb];
                flood_from(e, b);
            }
            ex = e->bp[2 * (e->n_bp - 1)]; ey = e->bp[2 * (e->n_bp - 1) + 1];
            e->n_bp--;
        }
        e->bp[2 * e->n_bp] = ex; e->bp[2 * e->n_bp + 1] = ey; e->n_bp++; /* :377 */
    }
    if (!e->alive) return e->cfg.death_penalty; /* :380-381 */
    if (e->n_burn == 0) {                        /* :384-387 */
        int healthy = 0, n = e->W * e->H;
        for (int i = 0; i < n; ++i) healthy += (e->type[i] == T_GRASS);
        double perc = (double)healthy / (double)(e->W * e->H);
        return e->cfg.contained_bonus * perc;
    }
    return e->cfg.default_reward; /* :390 */
}

/* World.get_state -- environment.py:399-402 */
void wfo_get_obs(const wfo_env* e, uint8_t* obs) {
    int n = e->W * e->H;
    for (int i = 0; i < n; ++i) {
        obs[3 * i + 0] = e->apos[i];
        obs[3 * i + 1] = e->type[i] == T_FIRE;
        obs[3 * i + 2] = !e->fm_inf[i];
    }
}

/* ForestFire.step -- forest_fire.py:30-49 */
int wfo_step(wfo_env* e, int action, uint8_t* obs, double* reward, int* done) {
    if (action >= 0 && action < 4) {
        if (!e->alive) return -1; /* agents[0] -> IndexError (Q8) */
        agent_move(e, action);
    }
    if (e->cfg.allow_dig_toggle && action == 4) {
        if (!e->alive) return -1;
        e->digging = !e->digging; /* Agent.toggle_digging environment.py:136-138 */
        agent_dig(e);
    }
    e->a_speed_iter -= 1;
    if (e->a_speed_iter == 0) {
        update(e);
        e->a_speed_iter = e->cfg.a_speed;
    }
    e->t += 1;
    if (obs) wfo_get_obs(e, obs);
    double r = get_reward(e);
    if (reward) *reward = r;
    if (done) *done = !e->running;
    return 0;
}

int wfo_stream_action(const wfo_env* e) {
    uint32_t w[4];
    stream_block(e, e->t >> 2, 1u, w); /* ACTION stream: step t = word (t & 3) of block t >> 2 */
    return (int)(w[e->t & 3u] % (uint32_t)e->cfg.n_actions);
}

/* DQN.choose_randomwalk_action(avoid_fire=True) -- DQN.py:353-389: walk clockwise round the fire
 * origin, re-drawing (at most 11 times) while the chosen move would step onto a burning cell.
 * np.random.choice(possible_actions) is draw j of the POLICY stream (stream 3) of step t. */
int wfo_policy_action(const wfo_env* e) {
    if (!e->alive) return 0; /* :356-357 */
    static const int DX[4] = {0, 0, 1, -1}, DY[4] = {-1, 1, 0, 0}; /* N S E W */
    const int mid_x = e->W / 2, mid_y = e->H / 2, ax = e->ax, ay = e->ay;
    int count = 0, action = 0, j = 0;
    for (;;) {
        int a0 = 0, a1 = 0; /* possible_actions, :369-376 */
        if (ax >= mid_x && ay > mid_y) { a0 = 1; a1 = 3; }  /* ["S", "W"] */
        if (ax > mid_x && ay <= mid_y) { a0 = 1; a1 = 2; }  /* ["S", "E"] */
        if (ax <= mid_x && ay < mid_y) { a0 = 0; a1 = 2; }  /* ["N", "E"] */
        if (ax < mid_x && ay >= mid_y) { a0 = 0; a1 = 3; }  /* ["N", "W"] */
        uint32_t w[4];
        stream_block(e, 3u * e->t + (uint32_t)(j >> 2), 3u, w);
        action = (w[j & 3] % 2u) ? a1 : a0; /* :379 */
        j++;
        int nx = ax + DX[action], ny = ay + DY[action];
        int fire_at_loc = inbounds(e, nx, ny) && is_burning(e, nx, ny); /* Agent.fire_in_direction environment.py:158-160 */
        if (!fire_at_loc || count > 10) break; /* :385-387 */
        count++;
    }
    return action;
}

void wfo_get_planes(const wfo_env* e, uint8_t* type, uint8_t* burning, uint8_t* fm_inf,
                    int32_t* fuel, double* temp, uint8_t* apos) {
    size_t n = (size_t)e->W * e->H;
    if (type) memcpy(type, e->type, n);
    if (burning) memcpy(burning, e->in_burn, n);
    if (fm_inf) memcpy(fm_inf, e->fm_inf, n);
    if (fuel) memcpy(fuel, e->fuel, n * sizeof(int32_t));
    if (temp) memcpy(temp, e->temp, n * sizeof(double));
    if (apos) memcpy(apos, e->apos, n);
}

void wfo_get_scalars(const wfo_env* e, int32_t out[16]) {
    memset(out, 0, 16 * sizeof(int32_t));
    out[0] = e->alive; out[1] = e->alive ? e->ax : -1; out[2] = e->alive ? e->ay : -1;
    out[3] = e->dead; out[4] = e->digging; out[5] = e->running; out[6] = e->fire_at_border;
    out[7] = e->n_bp; out[8] = (int32_t)e->episode; out[9] = (int32_t)e->t;
    out[10] = e->a_speed_iter; out[11] = e->wind_x; out[12] = e->wind_y; out[13] = e->n_burn;
}

double wfo_get_wind_speed(const wfo_env* e) { return e->wind_speed; }

/* METADATA['a_speed_iter'] is ONE process-wide counter in the reference (constants.py:41,
 * forest_fire.py:40-43).  A batch that steps in lockstep shares it; an env that sat out some
 * steps (finished, waiting for reset) must be re-synchronised by the harness. */
void wfo_set_a_speed_iter(wfo_env* e, int v) { e->a_speed_iter = v; }

void wfo_get_coef(const wfo_env* e, double coef[4]) {
    static const int DX[4] = {0, 0, 1, -1}, DY[4] = {-1, 1, 0, 0};
    for (int d = 0; d < 4; ++d) {
        int cx = DX[d], cy = DY[d], wx = e->wind_x, wy = e->wind_y;
        double angle = fabs(atan2((double)(wx * cy - wy * cx), (double)(wx * cx + wy * cy)));
        coef[d] = e->wind_speed * e->cfg.heat * pow(angle + 1.0, -1.0);
    }
}

void wfo_set_planes(wfo_env* e, const uint8_t* type, const uint8_t* burning, const uint8_t* fm_inf,
                    const int32_t* fuel, const double* temp) {
    size_t n = (size_t)e->W * e->H;
    if (type) memcpy(e->type, type, n);
    if (fm_inf) memcpy(e->fm_inf, fm_inf, n);
    if (fuel) memcpy(e->fuel, fuel, n * sizeof(int32_t));
    if (temp) memcpy(e->temp, temp, n * sizeof(double));
    if (burning) {
        e->n_burn = 0;
        for (size_t i = 0; i < n; ++i) {
            e->in_burn[i] = burning[i] != 0;
            if (burning[i]) e->burn_list[e->n_burn++] = (int32_t)i;
        }
    }
}

void wfo_set_agent(wfo_env* e, int alive, int ax, int ay, int visible, int dead, int digging) {
    memset(e->apos, 0, (size_t)e->W * e->H);
    e->alive = alive; e->ax = ax; e->ay = ay; e->dead = dead; e->digging = digging;
    if (alive && visible) e->apos[IDX(e, ax, ay)] = 1;
}

typedef struct {
    const wfo_config* cfg;
    int64_t env_id_base;
    int n_envs, n_steps, tid, n_threads;
    int64_t total, eps;
    double sum;
} rollout_job;

static void* rollout_worker(void* arg) {
    rollout_job* j = (rollout_job*)arg;
    const wfo_config* cfg = j->cfg;
    uint8_t* obs = (uint8_t*)malloc((size_t)cfg->width * cfg->height * 3);
    int64_t total = 0, eps = 0; /* thread-local accumulators (no false sharing on jobs[]) */
    double sum = 0.0;
    for (int i = j->tid; i < j->n_envs; i += j->n_threads) { /* envs are independent: static interleave */
        wfo_env* e = wfo_create(cfg, j->env_id_base + i);
        wfo_reset(e);
        for (int s = 0; s < j->n_steps; ++s) {
            double r; int d;
            wfo_step(e, wfo_stream_action(e), obs, &r, &d);
            sum += r + obs[(s * 7) % (cfg->width * cfg->height * 3)];
            total += 1;
            if (d) { wfo_reset(e); eps += 1; }
        }
        wfo_destroy(e);
    }
    free(obs);
    j->total = total; j->eps = eps; j->sum = sum;
    return NULL;
}

int64_t wfo_rollout(const wfo_config* cfg, int64_t env_id_base, int n_envs, int n_steps,
                    int n_threads, double* checksum, int64_t* episodes) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_envs) n_threads = n_envs;
    rollout_job* jobs = (rollout_job*)calloc((size_t)n_threads, sizeof(rollout_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].cfg = cfg; jobs[t].env_id_base = env_id_base; jobs[t].n_envs = n_envs;
        jobs[t].n_steps = n_steps; jobs[t].tid = t; jobs[t].n_threads = n_threads;
        pthread_create(&th[t], NULL, rollout_worker, &jobs[t]);
    }
    int64_t total = 0, eps = 0;
    double sum = 0.0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].total; eps += jobs[t].eps; sum += jobs[t].sum;
    }
    free(jobs); free(th);
    if (checksum) *checksum = sum;
    if (episodes) *episodes = eps;
    return total;
}

/* ---- persistent batch (bench.py --impl reference): same envs stepped call after call ---- */
struct wfo_batch {
    wfo_config cfg;
    int n_envs, n_threads;
    wfo_env** envs;
    uint8_t** obs; /* one scratch observation per thread */
};

typedef struct {
    struct wfo_batch* b;
    int tid, n_steps;
    int64_t total, eps;
    double sum;
} batch_job;

static void* batch_worker(void* arg) {
    batch_job* j = (batch_job*)arg;
    struct wfo_batch* b = j->b;
    int64_t total = 0, eps = 0;
    double sum = 0.0;
    uint8_t* obs = b->obs[j->tid];
    const int nb = b->cfg.width * b->cfg.height * 3;
    const int per = (b->n_envs + b->n_threads - 1) / b->n_threads; /* contiguous block per thread */
    const int lo = j->tid * per, hi = lo + per < b->n_envs ? lo + per : b->n_envs;
    for (int i = lo; i < hi; ++i) {
        wfo_env* e = b->envs[i];
        for (int s = 0; s < j->n_steps; ++s) {
            double r; int d;
            wfo_step(e, wfo_stream_action(e), obs, &r, &d);
            sum += r + obs[(s * 7 + i) % nb];
            total += 1;
            if (d) { wfo_reset(e); eps += 1; }
        }
    }
    j->total = total; j->eps = eps; j->sum = sum;
    return NULL;
}

struct wfo_batch* wfo_batch_create(const wfo_config* cfg, int64_t env_id_base, int n_envs, int n_threads) {
    struct wfo_batch* b = (struct wfo_batch*)calloc(1, sizeof(*b));
    b->cfg = *cfg;
    b->n_envs = n_envs;
    b->n_threads = n_threads < 1 ? 1 : (n_threads > n_envs ? n_envs : n_threads);
    b->envs = (wfo_env**)calloc((size_t)n_envs, sizeof(wfo_env*));
    b->obs = (uint8_t**)calloc((size_t)b->n_threads, sizeof(uint8_t*));
    for (int t = 0; t < b->n_threads; ++t) b->obs[t] = (uint8_t*)malloc((size_t)cfg->width * cfg->height * 3 + 64);
    for (int i = 0; i < n_envs; ++i) {
        b->envs[i] = wfo_create(cfg, env_id_base + i);
        wfo_reset(b->envs[i]);
    }
    return b;
}

void wfo_batch_destroy(struct wfo_batch* b) {
    if (!b) return;
    for (int i = 0; i < b->n_envs; ++i) wfo_destroy(b->envs[i]);
    for (int t = 0; t < b->n_threads; ++t) free(b->obs[t]);
    free(b->envs); free(b->obs); free(b);
}

/* n_steps ForestFire.step calls on every env (ACTION-stream actions, reset on done). */
int64_t wfo_batch_step(struct wfo_batch* b, int n_steps, double* checksum, int64_t* episodes) {
    batch_job* jobs = (batch_job*)calloc((size_t)b->n_threads, sizeof(batch_job));
    pthread_t* th = (pthread_t*)calloc((size_t)b->n_threads, sizeof(pthread_t));
    for (int t = 0; t < b->n_threads; ++t) {
        jobs[t].b = b; jobs[t].tid = t; jobs[t].n_steps = n_steps;
        if (b->n_threads > 1) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
        else batch_worker(&jobs[t]);
    }
    int64_t total = 0, eps = 0;
    double sum = 0.0;
    for (int t = 0; t < b->n_threads; ++t) {
        if (b->n_threads > 1) pthread_join(th[t], NULL);
        total += jobs[t].total; eps += jobs[t].eps; sum += jobs[t].sum;
    }
    free(jobs); free(th);
    if (checksum) *checksum += sum;
    if (episodes) *episodes += eps;
    return total;
}
