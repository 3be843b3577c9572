/*
 * TEST INFRASTRUCTURE ONLY — see foundation_oracle.h.
 *
 * Plain-C restatement of the reference's gather-trade-build step, written the way
 * the reference holds its state (float64 health maps, explicit order lists that
 * are stable-sorted every step, Python-float inventories).  It is deliberately
 * NOT structured like the CUDA product (bitfield cells, slot books, warp scans)
 * so that agreement between the two means something.
 *
 * Reference paths are relative to /root/reference/ai_economist/foundation/.
 */
#define _GNU_SOURCE /* pthread_setaffinity_np, sched_getaffinity (CPU baseline threads are pinned) */
#include "foundation_oracle.h"

#include <sched.h>

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define N_COMMODITIES 2 /* sorted collectible resources: Stone, Wood (continuous_double_auction.py:75-77) */
#define STONE 0
#define WOOD 1

typedef struct { int agent, price, life; } order_t;

typedef struct {
    /* Maps (base/world.py:36-112): one float64 health map per entity + House owner. */
    double *res[2];      /* Stone, Wood */
    double *src[2];      /* StoneSourceBlock, WoodSourceBlock */
    double *water;
    double *house;
    int16_t *owner;
    uint8_t *unoccupied;
    /* Agents (base/base_agent.py:62) */
    int *loc_r, *loc_c;
    double *coin, *esc_coin, *labor;
    double *inv[2], *esc[2];
    double *build_payment, *build_skill, *bonus_prob;
    /* ContinuousDoubleAuction state (continuous_double_auction.py:79-99) */
    order_t *bids[2], *asks[2];
    int n_bids[2], n_asks[2];
    int *n_orders[2];
    double *price_hist[2], *bid_hist[2], *ask_hist[2]; /* [A*P] */
    /* PeriodicBracketTax state (redistribution.py:319-330, 1106-1122) */
    int tax_pos;
    int rate_idx[ORC_MAX_BRACKETS];
    double curr_rates_obs[ORC_MAX_BRACKETS];
    double *last_coin, *last_income, *last_marg, *last_income_obs, *last_income_obs_sorted;
    double planner_mask_rates[ORC_MAX_RATES]; /* "new_taxes" mask for this episode */
    /* Episode logs behind get_metrics(), kept as running sums:
     *   Build.builds (build.py:150-159, 208-212): builds per agent
     *   ContinuousDoubleAuction.executed_trades (:316, 612-620): per (agent, commodity, seller/buyer) count + price sum
     *   PeriodicBracketTax._schedules / _occupancy / all_effective_tax_rates / total_collected_taxes / taxes
     *   (redistribution.py:862-905, 1149-1180) */
    double *n_builds;               /* [A] */
    double n_trades;
    double *trade_n, *trade_price;  /* [A][2 commodities][2: 0 as seller, 1 as buyer] */
    double tax_periods, tax_collected, tax_eff_sum;
    double tax_sched_sum[ORC_MAX_BRACKETS], tax_occupancy[ORC_MAX_BRACKETS];
    double *tax_income_pos, *tax_paid; /* [A] sums over tax days of max(0, income) and tax_paid */
    /* Scenario reward trackers (layout_from_file.py:160-163) */
    double *curr_metric; /* [A+1], planner last */
    int auto_warmup_integrator;
    int completions;
    int t;
    /* numpy legacy MT19937 (np.random global state, base_env.py:481-494) */
    uint32_t mt[624];
    int mt_pos;
    /* decoded actions for this step */
    int *act_build, *act_move, *act_buy[2], *act_sell[2];
    int act_tax[ORC_MAX_BRACKETS];
    /* outputs */
    float *a_map, *a_flat, *a_mask, *p_map, *p_flat, *p_agents, *p_mask;
    int16_t *a_idx, *p_idx;
    float time_obs;
    double *rew;
    int done;
} env_t;

struct orc_batch {
    orc_config cfg;
    orc_dims dims;
    int n_envs;
    env_t *envs;
    int has[ORC_MAX_COMP]; /* component kind present */
    /* action subspace tables in registration order (base_agent.py:97-169) */
    int n_sub;
    int sub_kind[8]; /* 0 build, 1 buy, 2 sell, 3 gather */
    int sub_c[8];
    int sub_n[8];
};

/* ------------------------------------------------------------------------- */
/* numpy legacy RandomState stream                                            */
/* ------------------------------------------------------------------------- */

/* numpy/random/src/mt19937/mt19937.c: mt19937_gen (standard MT19937 twist). */
static void mt_gen(env_t *s) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
    uint32_t y;
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
        y = (s->mt[kk] & UPPER) | (s->mt[kk + 1] & LOWER);
        s->mt[kk] = s->mt[kk + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    for (; kk < 623; kk++) {
        y = (s->mt[kk] & UPPER) | (s->mt[kk + 1] & LOWER);
        s->mt[kk] = s->mt[kk + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    y = (s->mt[623] & UPPER) | (s->mt[0] & LOWER);
    s->mt[623] = s->mt[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    s->mt_pos = 0;
}

static uint32_t mt_next32(env_t *s) {
    uint32_t y;
    if (s->mt_pos == 624) mt_gen(s);
    y = s->mt[s->mt_pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* mt19937_next_double: 53-bit double from two 32-bit words; this is np.random.rand(). */
static double np_rand(env_t *s) {
    int32_t a = mt_next32(s) >> 5, b = mt_next32(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* legacy random_interval(max): masked rejection on 32-bit draws (max <= 0xffffffff). */
static uint32_t np_interval(env_t *s, uint32_t max) {
    uint32_t mask = max, value;
    if (max == 0) return 0;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    while ((value = (mt_next32(s) & mask)) > max) {}
    return value;
}

/* np.random.permutation(n): arange then RandomState.shuffle (Fisher-Yates from the top).
 * Used by World.get_random_order_agents, base/world.py:418-422. */
static void np_permutation(env_t *s, int n, int *out) {
    int i;
    for (i = 0; i < n; i++) out[i] = i;
    for (i = n - 1; i >= 1; i--) {
        int j = (int)np_interval(s, (uint32_t)i);
        int tmp = out[j]; out[j] = out[i]; out[i] = tmp;
    }
}

/* ------------------------------------------------------------------------- */
/* dims                                                                       */
/* ------------------------------------------------------------------------- */

static int cfg_has(const orc_config *cfg, int kind) {
    int i;
    for (i = 0; i < cfg->n_comp; i++) if (cfg->comp[i] == kind) return 1;
    return 0;
}

static int planner_has_tax_actions(const orc_config *cfg) {
    /* redistribution.py:929-939 */
    return cfg_has(cfg, ORC_COMP_TAX) && cfg->tax_model == ORC_TAX_MODEL_WRAPPER && !cfg->disable_taxes;
}

int orc_dims_from_config(const orc_config *cfg, orc_dims *d) {
    int A = cfg->n_agents, P = cfg->max_bid_ask + 1, C = N_COMMODITIES;
    int has_build = cfg_has(cfg, ORC_COMP_BUILD), has_cda = cfg_has(cfg, ORC_COMP_CDA);
    int has_gather = cfg_has(cfg, ORC_COMP_GATHER), has_tax = cfg_has(cfg, ORC_COMP_TAX);
    int B = cfg->n_brackets;
    int n_sub = 0, n_single = 0;
    memset(d, 0, sizeof(*d));
    d->n_map_ch = cfg->has_water ? 6 : 5;
    d->win = 2 * cfg->obs_range + 1;
    /* agent flat: see SURVEY 8(a) row O2 / base_env.py:562-612 */
    d->flat_a = (has_build ? 2 : 0) + (has_cda ? C * (5 * P + 1) : 0) + (has_gather ? 1 : 0) +
                (has_tax ? (B + 2 + A + 1 + 1) : 0) + 1 /* time */ + 3 /* inventory */ +
                (cfg->full_observability ? 0 : 2) /* loc */;
    d->flat_p = (has_cda ? C * (3 * P + 1) : 0) + (has_tax ? (B + 2 + A + 1) : 0) + 1 + 3;
    d->flat_pa = (has_tax ? 3 : 0) + (cfg->full_observability ? 0 : 3 + (cfg->planner_gets_spatial_info ? 2 : 0));
    d->a_map_elems = cfg->full_observability ? d->n_map_ch * cfg->height * cfg->width : (d->n_map_ch + 1) * d->win * d->win;
    d->a_idx_elems = cfg->full_observability ? 2 * cfg->height * cfg->width : 2 * d->win * d->win;
    {
        int i;
        for (i = 0; i < cfg->n_comp; i++) {
            if (cfg->comp[i] == ORC_COMP_BUILD) { n_sub += 1; n_single += 1; }
            if (cfg->comp[i] == ORC_COMP_CDA) { n_sub += 2 * C; n_single += 2 * C * P; }
            if (cfg->comp[i] == ORC_COMP_GATHER) { n_sub += 1; n_single += 4; }
        }
    }
    /* base_agent.py:440-460 */
    d->mask_a = cfg->multi_action_agents ? (n_single + n_sub) : (1 + n_single);
    d->n_act_a = cfg->multi_action_agents ? n_sub : 1;
    if (planner_has_tax_actions(cfg) && cfg->single_action_planner) {
        d->n_act_p = 1;  /* one index into [NO-OP] ++ bracket 0 rates ++ bracket 1 rates ... (base_agent.py:109-114) */
        d->mask_p = 1 + B * cfg->n_disc_rates;
    } else if (planner_has_tax_actions(cfg)) {
        d->n_act_p = B;
        d->mask_p = B * (1 + cfg->n_disc_rates);
    } else {
        d->n_act_p = 0;
        d->mask_p = 1; /* passive multi-action agent: [1] (base_agent.py:447-448) */
    }
    d->book_cap = A * cfg->max_num_orders;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* create / destroy                                                           */
/* ------------------------------------------------------------------------- */

static void *zalloc(size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) abort(); return p; }

orc_batch *orc_create(const orc_config *cfg, int32_t n_envs) {
    orc_batch *b = (orc_batch *)zalloc(sizeof(*b));
    int A = cfg->n_agents, HW = cfg->height * cfg->width, P = cfg->max_bid_ask + 1;
    int e, c, i;
    b->cfg = *cfg;
    orc_dims_from_config(cfg, &b->dims);
    b->n_envs = n_envs;
    for (i = 0; i < ORC_MAX_COMP; i++) b->has[i] = cfg_has(cfg, i);
    /* action subspaces in component-list order (base_agent.py:124-169) */
    for (i = 0; i < cfg->n_comp; i++) {
        if (cfg->comp[i] == ORC_COMP_BUILD) {
            b->sub_kind[b->n_sub] = 0; b->sub_c[b->n_sub] = 0; b->sub_n[b->n_sub++] = 1;
        } else if (cfg->comp[i] == ORC_COMP_CDA) {
            for (c = 0; c < N_COMMODITIES; c++) { /* Buy_c then Sell_c, continuous_double_auction.py:418-428 */
                b->sub_kind[b->n_sub] = 1; b->sub_c[b->n_sub] = c; b->sub_n[b->n_sub++] = P;
                b->sub_kind[b->n_sub] = 2; b->sub_c[b->n_sub] = c; b->sub_n[b->n_sub++] = P;
            }
        } else if (cfg->comp[i] == ORC_COMP_GATHER) {
            b->sub_kind[b->n_sub] = 3; b->sub_c[b->n_sub] = 0; b->sub_n[b->n_sub++] = 4;
        }
    }
    b->envs = (env_t *)zalloc(sizeof(env_t) * n_envs);
    for (e = 0; e < n_envs; e++) {
        env_t *s = &b->envs[e];
        const orc_dims *d = &b->dims;
        for (c = 0; c < 2; c++) {
            s->res[c] = (double *)zalloc(sizeof(double) * HW);
            s->src[c] = (double *)zalloc(sizeof(double) * HW);
            s->inv[c] = (double *)zalloc(sizeof(double) * A);
            s->esc[c] = (double *)zalloc(sizeof(double) * A);
            s->bids[c] = (order_t *)zalloc(sizeof(order_t) * (d->book_cap + A));
            s->asks[c] = (order_t *)zalloc(sizeof(order_t) * (d->book_cap + A));
            s->n_orders[c] = (int *)zalloc(sizeof(int) * A);
            s->price_hist[c] = (double *)zalloc(sizeof(double) * A * P);
            s->bid_hist[c] = (double *)zalloc(sizeof(double) * A * P);
            s->ask_hist[c] = (double *)zalloc(sizeof(double) * A * P);
            s->act_buy[c] = (int *)zalloc(sizeof(int) * A);
            s->act_sell[c] = (int *)zalloc(sizeof(int) * A);
        }
        s->water = (double *)zalloc(sizeof(double) * HW);
        s->house = (double *)zalloc(sizeof(double) * HW);
        s->owner = (int16_t *)zalloc(sizeof(int16_t) * HW);
        s->unoccupied = (uint8_t *)zalloc(HW);
        s->loc_r = (int *)zalloc(sizeof(int) * A);
        s->loc_c = (int *)zalloc(sizeof(int) * A);
        s->coin = (double *)zalloc(sizeof(double) * A);
        s->esc_coin = (double *)zalloc(sizeof(double) * A);
        s->labor = (double *)zalloc(sizeof(double) * A);
        s->build_payment = (double *)zalloc(sizeof(double) * A);
        s->build_skill = (double *)zalloc(sizeof(double) * A);
        s->bonus_prob = (double *)zalloc(sizeof(double) * A);
        s->last_coin = (double *)zalloc(sizeof(double) * A);
        s->last_income = (double *)zalloc(sizeof(double) * A);
        s->last_marg = (double *)zalloc(sizeof(double) * A);
        s->last_income_obs = (double *)zalloc(sizeof(double) * A);
        s->last_income_obs_sorted = (double *)zalloc(sizeof(double) * A);
        s->curr_metric = (double *)zalloc(sizeof(double) * (A + 1));
        s->n_builds = (double *)zalloc(sizeof(double) * A);
        s->trade_n = (double *)zalloc(sizeof(double) * A * 4);
        s->trade_price = (double *)zalloc(sizeof(double) * A * 4);
        s->tax_income_pos = (double *)zalloc(sizeof(double) * A);
        s->tax_paid = (double *)zalloc(sizeof(double) * A);
        s->act_build = (int *)zalloc(sizeof(int) * A);
        s->act_move = (int *)zalloc(sizeof(int) * A);
        s->a_map = (float *)zalloc(sizeof(float) * A * d->a_map_elems);
        s->a_idx = (int16_t *)zalloc(sizeof(int16_t) * A * d->a_idx_elems);
        s->a_flat = (float *)zalloc(sizeof(float) * A * d->flat_a);
        s->a_mask = (float *)zalloc(sizeof(float) * A * d->mask_a);
        s->p_map = (float *)zalloc(sizeof(float) * d->n_map_ch * HW);
        s->p_idx = (int16_t *)zalloc(sizeof(int16_t) * 2 * HW);
        s->p_flat = (float *)zalloc(sizeof(float) * d->flat_p);
        s->p_agents = (float *)zalloc(sizeof(float) * A * (d->flat_pa > 0 ? d->flat_pa : 1));
        s->p_mask = (float *)zalloc(sizeof(float) * d->mask_p);
        s->rew = (double *)zalloc(sizeof(double) * (A + 1));
    }
    return b;
}

void orc_destroy(orc_batch *b) {
    int e, c;
    if (!b) return;
    for (e = 0; e < b->n_envs; e++) {
        env_t *s = &b->envs[e];
        for (c = 0; c < 2; c++) {
            free(s->res[c]); free(s->src[c]); free(s->inv[c]); free(s->esc[c]);
            free(s->bids[c]); free(s->asks[c]); free(s->n_orders[c]);
            free(s->price_hist[c]); free(s->bid_hist[c]); free(s->ask_hist[c]);
            free(s->act_buy[c]); free(s->act_sell[c]);
        }
        free(s->water); free(s->house); free(s->owner); free(s->unoccupied);
        free(s->loc_r); free(s->loc_c); free(s->coin); free(s->esc_coin); free(s->labor);
        free(s->build_payment); free(s->build_skill); free(s->bonus_prob);
        free(s->last_coin); free(s->last_income); free(s->last_marg);
        free(s->last_income_obs); free(s->last_income_obs_sorted);
        free(s->curr_metric); free(s->act_build); free(s->act_move);
        free(s->n_builds); free(s->trade_n); free(s->trade_price); free(s->tax_income_pos); free(s->tax_paid);
        free(s->a_map); free(s->a_idx); free(s->a_flat); free(s->a_mask);
        free(s->p_map); free(s->p_idx); free(s->p_flat); free(s->p_agents); free(s->p_mask);
        free(s->rew);
    }
    free(b->envs);
    free(b);
}

/* ------------------------------------------------------------------------- */
/* world helpers (base/world.py)                                              */
/* ------------------------------------------------------------------------- */

/* Maps.accessibility: AND over blocking (Water == 0) and private (House owner in {-1, self}),
 * world.py:213-228, 256-259, 300-305. */
static int accessible(const orc_batch *b, const env_t *s, int r, int c, int agent) {
    int W = b->cfg.width, k = r * W + c;
    if (s->water[k] != 0) return 0;
    if (!(s->owner[k] == -1 || s->owner[k] == agent)) return 0;
    return 1;
}

/* World.can_agent_occupy, world.py:424-440 */
static int can_occupy(const orc_batch *b, const env_t *s, int r, int c, int agent) {
    if (!(r >= 0 && r < b->cfg.height && c >= 0 && c < b->cfg.width)) return 0;
    if (!accessible(b, s, r, c, agent)) return 0;
    return s->unoccupied[r * b->cfg.width + c] ? 1 : 0;
}

/* ------------------------------------------------------------------------- */
/* Build  (components/build.py)                                               */
/* ------------------------------------------------------------------------- */

/* Build.agent_can_build, build.py:70-83 (+ world.py:284-293 for the location queries) */
static int agent_can_build(const orc_batch *b, const env_t *s, int a) {
    int k = s->loc_r[a] * b->cfg.width + s->loc_c[a];
    if (s->inv[WOOD][a] < 1) return 0;
    if (s->inv[STONE][a] < 1) return 0;
    /* location_resources: any collectible with health > 0 */
    if (s->res[STONE][k] > 0 || s->res[WOOD][k] > 0) return 0;
    /* location_landmarks: any non-resource map with value > 0 (House health, Water, SourceBlocks) */
    if (s->house[k] > 0 || s->water[k] > 0 || s->src[STONE][k] > 0 || s->src[WOOD][k] > 0) return 0;
    return 1;
}

/* Build.component_step, build.py:112-161 */
static void build_step(const orc_batch *b, env_t *s) {
    int A = b->cfg.n_agents, i;
    int *order = (int *)alloca(sizeof(int) * A);
    np_permutation(s, A, order); /* world.get_random_order_agents() */
    for (i = 0; i < A; i++) {
        int a = order[i];
        if (s->act_build[a] == 1 && agent_can_build(b, s, a)) {
            int k = s->loc_r[a] * b->cfg.width + s->loc_c[a];
            s->inv[WOOD][a] -= 1;
            s->inv[STONE][a] -= 1;
            s->house[k] = 1; /* world.create_landmark -> Maps.set_point, world.py:240-259 */
            s->owner[k] = (int16_t)a;
            s->coin[a] += s->build_payment[a];
            s->labor[a] += b->cfg.build_labor;
            s->n_builds[a] += 1; /* self.builds[-1].append({builder, loc, income}) :150-159 */
        }
    }
}

/* ------------------------------------------------------------------------- */
/* ContinuousDoubleAuction (components/continuous_double_auction.py)          */
/* ------------------------------------------------------------------------- */

/* Python's sorted() is stable, also with reverse=True.  Insertion sort is stable. */
static void sort_bids(order_t *v, int n) { /* key (bid, bid_lifetime), reverse=True: :246-250 */
    int i, j;
    for (i = 1; i < n; i++) {
        order_t x = v[i];
        for (j = i - 1; j >= 0; j--) {
            int lt = (v[j].price < x.price) || (v[j].price == x.price && v[j].life < x.life);
            if (!lt) break; /* keep earlier element first unless strictly smaller key */
            v[j + 1] = v[j];
        }
        v[j + 1] = x;
    }
}
static void sort_asks(order_t *v, int n) { /* key (ask, -ask_lifetime) ascending: :251-253 */
    int i, j;
    for (i = 1; i < n; i++) {
        order_t x = v[i];
        for (j = i - 1; j >= 0; j--) {
            int gt = (v[j].price > x.price) || (v[j].price == x.price && -v[j].life > -x.life);
            if (!gt) break;
            v[j + 1] = v[j];
        }
        v[j + 1] = x;
    }
}
static void remove_at(order_t *v, int *n, int idx) {
    memmove(&v[idx], &v[idx + 1], sizeof(order_t) * (*n - idx - 1));
    (*n)--;
}

/* create_bid :168-198 */
static void create_bid(const orc_batch *b, env_t *s, int c, int a, int max_payment) {
    int P = b->cfg.max_bid_ask + 1;
    if (!(s->n_orders[c][a] < b->cfg.max_num_orders) || s->coin[a] < max_payment) return;
    s->bids[c][s->n_bids[c]].agent = a;
    s->bids[c][s->n_bids[c]].price = max_payment;
    s->bids[c][s->n_bids[c]].life = 0;
    s->n_bids[c]++;
    s->bid_hist[c][a * P + max_payment] += 1;
    s->n_orders[c][a] += 1;
    { /* inventory_to_escrow, base_agent.py:279-298 */
        double tr = fmin(s->coin[a], (double)max_payment);
        s->coin[a] -= tr;
        s->esc_coin[a] += tr;
    }
    s->labor[a] += b->cfg.order_labor;
}

/* create_ask :200-229 */
static void create_ask(const orc_batch *b, env_t *s, int c, int a, int min_income) {
    int P = b->cfg.max_bid_ask + 1;
    if (!(s->n_orders[c][a] < b->cfg.max_num_orders && s->inv[c][a] > 0)) return;
    s->asks[c][s->n_asks[c]].agent = a;
    s->asks[c][s->n_asks[c]].price = min_income;
    s->asks[c][s->n_asks[c]].life = 0;
    s->n_asks[c]++;
    s->ask_hist[c][a * P + min_income] += 1;
    s->n_orders[c][a] += 1;
    {
        double tr = fmin(s->inv[c][a], 1.0);
        s->inv[c][a] -= tr;
        s->esc[c][a] += tr;
    }
    s->labor[a] += b->cfg.order_labor;
}

/* match_orders :231-350 */
static void match_orders(const orc_batch *b, env_t *s) {
    int A = b->cfg.n_agents, P = b->cfg.max_bid_ask + 1, c, i;
    int *possible = (int *)alloca(sizeof(int) * A);
    for (c = 0; c < N_COMMODITIES; c++) {
        order_t *bids = s->bids[c], *asks = s->asks[c];
        int keep_checking = 1;
        for (i = 0; i < A; i++) possible[i] = 1;
        sort_bids(bids, s->n_bids[c]);
        sort_asks(asks, s->n_asks[c]);
        for (;;) {
            int any = 0, idx_bid = 0, idx_ask = 0;
            for (i = 0; i < A; i++) any |= possible[i];
            if (!(any && keep_checking)) break;
            for (;;) {
                if (idx_bid >= s->n_bids[c]) { keep_checking = 0; break; }
                if (!possible[bids[idx_bid].agent]) { idx_bid++; }
                else if (idx_ask >= s->n_asks[c]) { possible[bids[idx_bid].agent] = 0; break; }
                else if (asks[idx_ask].agent == bids[idx_bid].agent) { idx_ask++; }
                else if (bids[idx_bid].price < asks[idx_ask].price) { possible[bids[idx_bid].agent] = 0; break; }
                else {
                    order_t bid = bids[idx_bid], ask = asks[idx_ask];
                    int price, buyer = bid.agent, seller = ask.agent;
                    remove_at(bids, &s->n_bids[c], idx_bid);
                    remove_at(asks, &s->n_asks[c], idx_ask);
                    price = (bid.life <= ask.life) ? ask.price : bid.price; /* :297-304 */
                    s->bid_hist[c][buyer * P + bid.price] -= 1;
                    s->ask_hist[c][seller * P + ask.price] -= 1;
                    s->n_orders[c][seller] -= 1;
                    s->n_orders[c][buyer] -= 1;
                    s->price_hist[c][seller * P + price] += 1;
                    s->esc[c][seller] -= 1;
                    s->inv[c][buyer] += 1;
                    s->esc_coin[buyer] -= bid.price;
                    s->coin[seller] += price;
                    s->coin[buyer] += bid.price - price;
                    /* executed_trades[-1].append(trade), cost == income == price :305-316 */
                    s->n_trades += 1;
                    s->trade_n[(seller * 2 + c) * 2 + 0] += 1; s->trade_price[(seller * 2 + c) * 2 + 0] += price;
                    s->trade_n[(buyer * 2 + c) * 2 + 1] += 1;  s->trade_price[(buyer * 2 + c) * 2 + 1] += price;
                    break;
                }
            }
        }
    }
}

/* remove_expired_orders :352-406 */
static void remove_expired(const orc_batch *b, env_t *s) {
    int P = b->cfg.max_bid_ask + 1, D = b->cfg.order_duration, c, i, n;
    for (c = 0; c < N_COMMODITIES; c++) {
        n = 0;
        for (i = 0; i < s->n_bids[c]; i++) {
            order_t o = s->bids[c][i];
            o.life += 1;
            if (o.life <= D) s->bids[c][n++] = o;
            else {
                double tr = fmin(s->esc_coin[o.agent], (double)o.price); /* escrow_to_inventory */
                s->esc_coin[o.agent] -= tr;
                s->coin[o.agent] += tr;
                s->bid_hist[c][o.agent * P + o.price] -= 1;
                s->n_orders[c][o.agent] -= 1;
            }
        }
        s->n_bids[c] = n;
        n = 0;
        for (i = 0; i < s->n_asks[c]; i++) {
            order_t o = s->asks[c][i];
            o.life += 1;
            if (o.life <= D) s->asks[c][n++] = o;
            else {
                double tr = fmin(s->esc[c][o.agent], 1.0);
                s->esc[c][o.agent] -= tr;
                s->inv[c][o.agent] += tr;
                s->ask_hist[c][o.agent * P + o.price] -= 1;
                s->n_orders[c][o.agent] -= 1;
            }
        }
        s->n_asks[c] = n;
    }
}

/* component_step :440-489 */
static void cda_step(const orc_batch *b, env_t *s) {
    int A = b->cfg.n_agents, P = b->cfg.max_bid_ask + 1, c, a, p;
    for (c = 0; c < N_COMMODITIES; c++) {
        for (a = 0; a < A; a++) {
            for (p = 0; p < P; p++) s->price_hist[c][a * P + p] *= 0.995;
            if (s->act_buy[c][a] != 0) create_bid(b, s, c, a, s->act_buy[c][a] - 1);
            if (s->act_sell[c][a] != 0) create_ask(b, s, c, a, s->act_sell[c][a] - 1);
        }
    }
    match_orders(b, s);
    remove_expired(b, s);
}

/* ------------------------------------------------------------------------- */
/* Gather (components/move.py)                                                */
/* ------------------------------------------------------------------------- */

/* component_step, move.py:93-153 */
static void gather_step(const orc_batch *b, env_t *s) {
    int A = b->cfg.n_agents, W = b->cfg.width, i, c;
    int *order = (int *)alloca(sizeof(int) * A);
    np_permutation(s, A, order);
    for (i = 0; i < A; i++) {
        int a = order[i], action = s->act_move[a];
        int r = s->loc_r[a], cc = s->loc_c[a], nr = r, nc = cc;
        if (action != 0) {
            if (action == 1) { nr = r; nc = cc - 1; }       /* Left */
            else if (action == 2) { nr = r; nc = cc + 1; }  /* Right */
            else if (action == 3) { nr = r - 1; nc = cc; }  /* Up */
            else { nr = r + 1; nc = cc; }                   /* Down */
            /* world.set_agent_loc, world.py:454-460 + Maps.set_agent_loc :150-173 */
            if (can_occupy(b, s, nr, nc, a)) {
                s->unoccupied[r * W + cc] = 1;
                s->loc_r[a] = nr; s->loc_c[a] = nc;
                s->unoccupied[nr * W + nc] = 0;
            }
            nr = s->loc_r[a]; nc = s->loc_c[a];
            if (nr != r || nc != cc) s->labor[a] += b->cfg.move_labor;
        }
        /* world.location_resources(new_r, new_c): resources with health > 0, in _resources
         * order Stone, Wood (world.py:284-288); collected if health >= 1 (move.py:136). */
        for (c = 0; c < N_COMMODITIES; c++) {
            int k = nr * W + nc;
            if (s->res[c][k] > 0 && s->res[c][k] >= 1) {
                int n_gathered = 1 + (np_rand(s) < s->bonus_prob[a] ? 1 : 0);
                s->inv[c][a] += n_gathered;
                s->res[c][k] = fmax(0.0, s->res[c][k] - 1); /* consume_resource, world.py:481-483 */
                s->labor[a] += b->cfg.collect_labor;
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* PeriodicBracketTax (components/redistribution.py)                          */
/* ------------------------------------------------------------------------- */

/* curr_marginal_rates :381-405 */
static void curr_marginal_rates(const orc_batch *b, const env_t *s, double *out) {
    int i;
    for (i = 0; i < b->cfg.n_brackets; i++) {
        if (b->cfg.tax_model == ORC_TAX_MODEL_WRAPPER) out[i] = b->cfg.disc_rates[s->rate_idx[i]];
        else {
            out[i] = b->cfg.fixed_rates[i];
            if (b->cfg.tax_annealing) { /* np.minimum(schedule, curr_rate_max), :390-394, :400-413; utils.py:10-57 */
                /* the limit is refreshed in generate_masks, i.e. after the observations of a reset were built: the reset
                 * observation (t == 0) still uses the previous episode's limit */
                int done_eps = (s->t == 0 && s->completions > 0) ? s->completions - 1 : s->completions;
                double vis = fmax(0.0, fmin(1.0, b->cfg.annealing_slope * ((double)done_eps - b->cfg.annealing_warmup)));
                out[i] = fmin(out[i], vis * b->cfg.rate_max);
            }
        }
    }
}

/* marginal_rate :837-844 */
static double marginal_rate(const orc_batch *b, const env_t *s, double income) {
    int B = b->cfg.n_brackets, i, arg = 0, found = 0;
    double rates[ORC_MAX_BRACKETS];
    if (income < 0) return 0.0;
    curr_marginal_rates(b, s, rates);
    for (i = 0; i < B; i++) {
        double lo = b->cfg.bracket_cutoffs[i];
        double hi = (i + 1 < B) ? b->cfg.bracket_cutoffs[i + 1] : INFINITY;
        if (income >= lo && income < hi) { if (!found) { arg = i; found = 1; } }
    }
    return rates[arg]; /* np.argmax of an all-False array is 0 */
}

/* taxes_due :846-851 */
static double taxes_due(const orc_batch *b, const env_t *s, double income) {
    int B = b->cfg.n_brackets, i;
    double rates[ORC_MAX_BRACKETS], sum = 0.0;
    curr_marginal_rates(b, s, rates);
    for (i = 0; i < B; i++) {
        double size = ((i + 1 < B) ? b->cfg.bracket_cutoffs[i + 1] : INFINITY) - b->cfg.bracket_cutoffs[i];
        double past = fmax(0.0, income - b->cfg.bracket_cutoffs[i]);
        double bin_income = fmin(size, past);
        sum += rates[i] * bin_income;
    }
    return sum;
}

static int cmp_double_idx(const void *x, const void *y) {
    double a = *(const double *)x, c = *(const double *)y;
    return (a > c) - (a < c);
}

/* enact_taxes :853-915 */
static void enact_taxes(const orc_batch *b, env_t *s) {
    int A = b->cfg.n_agents, B = b->cfg.n_brackets, a, i;
    double net = 0.0, lump, rates[ORC_MAX_BRACKETS];
    curr_marginal_rates(b, s, rates);
    for (i = 0; i < B; i++) s->tax_sched_sum[i] += rates[i]; /* _schedules[k].append(rate) :862-867 */
    s->tax_periods += 1;
    for (a = 0; a < A; a++) {
        double income = (s->coin[a] + s->esc_coin[a]) - s->last_coin[a];
        double due = taxes_due(b, s, income);
        double eff = fmin(s->coin[a], due); /* don't take from escrow */
        double marg = marginal_rate(b, s, income);
        int bin = 0, found = 0;
        s->coin[a] -= eff;
        net += eff;
        s->last_income[a] = income;
        s->last_marg[a] = marg;
        s->tax_eff_sum += eff / fmax(0.000001, income); /* all_effective_tax_rates.append :880, 894 */
        if (income >= 0) /* income_bin :828-835; negative income -> cutoff 0 = bracket 0 */
            for (i = 0; i < B; i++) {
                double lo = b->cfg.bracket_cutoffs[i], hi = (i + 1 < B) ? b->cfg.bracket_cutoffs[i + 1] : INFINITY;
                if (income >= lo && income < hi && !found) { bin = i; found = 1; }
            }
        s->tax_occupancy[bin] += 1;               /* :895 */
        s->tax_income_pos[a] += fmax(0.0, income); /* :1170-1175 via self.taxes */
        s->tax_paid[a] += eff;
    }
    s->tax_collected += net; /* :897 */
    lump = net / A;
    for (a = 0; a < A; a++) {
        s->coin[a] += lump;
        s->last_coin[a] = s->coin[a] + s->esc_coin[a];
    }
    for (a = 0; a < A; a++) s->last_income_obs[a] = s->last_income[a] / b->cfg.period;
    memcpy(s->last_income_obs_sorted, s->last_income_obs, sizeof(double) * A);
    qsort(s->last_income_obs_sorted, A, sizeof(double), cmp_double_idx); /* values only: order of ties is irrelevant */
}

/* component_step :945-972 */
static void tax_step(const orc_batch *b, env_t *s) {
    int B = b->cfg.n_brackets, i;
    if (s->tax_pos == 1) {
        if (b->cfg.tax_model == ORC_TAX_MODEL_WRAPPER && !b->cfg.disable_taxes) {
            for (i = 0; i < B; i++) { /* set_new_period_rates_model :419-434 */
                int act = s->act_tax[i];
                if (act != 0) s->rate_idx[i] = act - 1;
            }
        }
        curr_marginal_rates(b, s, s->curr_rates_obs);
    }
    if (s->tax_pos >= b->cfg.period) {
        enact_taxes(b, s);
        s->tax_pos = 0;
    }
    s->tax_pos += 1;
}

/* numpy's pairwise summation of a contiguous float64 array (np.sum; numpy/core/src/umath/loops_utils.h.src
 * pairwise_sum, blocks < 128 elements: 8 accumulators, then the remainder) */
static double np_sum(const double *a, int n) {
    double r[8], res;
    int i, j;
    if (n < 8) { res = 0.0; for (i = 0; i < n; i++) res += a[i]; return res; }
    for (j = 0; j < 8; j++) r[j] = a[j];
    for (i = 8; i < n - (n % 8); i += 8) for (j = 0; j < 8; j++) r[j] += a[i + j];
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

/* WealthRedistribution.component_step, redistribution.py:52-68 */
static void wealth_step(const orc_batch *b, env_t *s) {
    int A = b->cfg.n_agents, a;
    double *tot = (double *)alloca(sizeof(double) * A), target_share;
    for (a = 0; a < A; a++) tot[a] = s->coin[a] + s->esc_coin[a]; /* ic + ec */
    target_share = np_sum(tot, A) / A;
    for (a = 0; a < A; a++) s->coin[a] = target_share - s->esc_coin[a];
}

/* additional_reset_steps :1106-1139 */
static void tax_reset(const orc_batch *b, env_t *s) {
    int A = b->cfg.n_agents, a, i;
    for (i = 0; i < ORC_MAX_BRACKETS; i++) s->rate_idx[i] = 0;
    s->tax_pos = 1;
    for (a = 0; a < A; a++) {
        s->last_coin[a] = s->coin[a] + s->esc_coin[a];
        s->last_income[a] = 0; s->last_marg[a] = 0;
        s->last_income_obs[a] = 0; s->last_income_obs_sorted[a] = 0;
    }
    curr_marginal_rates(b, s, s->curr_rates_obs);
}

/* ------------------------------------------------------------------------- */
/* scenario_step: resource regeneration (layout_from_file.py:372-410,         */
/* identical copy dynamic_layout.py:433-471); max_health == 1.               */
/* ------------------------------------------------------------------------- */
/* scipy.signal.convolve2d(health, kernel, "same") at one cell, zero fill outside the map: a running float64 sum of
 * health * kernel_value over the d x d window (every kernel entry is regen_weight / d^2, so the term order only
 * matters through the number of non-zero terms). */
static double regen_prob(const orc_batch *b, const env_t *s, int c, int k) {
    int H = b->cfg.height, W = b->cfg.width, hw = b->cfg.regen_halfwidth[c], d = 1 + 2 * hw;
    double kv = (b->cfg.regen_weight[c] * 1.0) / (double)(d * d), p = 0.0;
    int r0 = k / W, c0 = k % W, dr, dc;
    for (dr = -hw; dr <= hw; dr++)
        for (dc = -hw; dc <= hw; dc++) {
            int r = r0 + dr, cc = c0 + dc;
            if (r < 0 || r >= H || cc < 0 || cc >= W) continue;
            p += fmax(s->res[c][r * W + cc], s->src[c][r * W + cc]) * kv;
        }
    return p;
}

static void scenario_step(const orc_batch *b, env_t *s) {
    int HW = b->cfg.height * b->cfg.width, k, ri;
    static const int order[2] = {WOOD, STONE}; /* resources = ["Wood", "Stone"] */
    for (ri = 0; ri < 2; ri++) {
        int c = order[ri];
        double w = b->cfg.regen_weight[c];
        /* the convolution reads the pre-update health map; with max_health == 1 health == source map, which this
         * loop never changes, so evaluating it cell by cell is the same */
        for (k = 0; k < HW; k++) {
            double health = fmax(s->res[c][k], s->src[c][k]);
            double u = np_rand(s); /* np.random.rand(*health.shape): one draw per cell, row-major */
            int respawn = b->cfg.regen_halfwidth[c] ? (u < regen_prob(b, s, c, k))
                                                    : (u < (health * w)); /* 1x1 kernel of value regen_weight */
            int spawnable;
            {   /* maps.empty (all maps sum == 0) + resource + source > 0, then *= source > 0 */
                double sum = s->res[STONE][k] + s->res[WOOD][k] + s->house[k] + s->water[k] +
                             s->src[STONE][k] + s->src[WOOD][k];
                spawnable = (((sum == 0) + s->res[c][k] + s->src[c][k]) > 0) && (s->src[c][k] > 0);
            }
            respawn = respawn && spawnable;
            s->res[c][k] = fmin(s->res[c][k] + respawn, 1.0);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* rewards (scenarios/utils/rewards.py, social_metrics.py,                    */
/* layout_from_file.py:249-318, 519-559)                                      */
/* ------------------------------------------------------------------------- */

static double energy_weight(const orc_batch *b, const env_t *s) { /* layout_from_file.py:249-267 */
    if (b->cfg.energy_warmup_constant <= 0.0) return 1.0;
    if (!b->cfg.energy_warmup_auto) return 1.0 - exp(-(double)s->completions / b->cfg.energy_warmup_constant);
    return 1.0 - exp(-(double)s->auto_warmup_integrator / b->cfg.energy_warmup_constant);
}

static double get_gini(const double *x, int n) { /* social_metrics.py:10-46 */
    int i, j;
    if (n < 30) {
        double diff = 0.0, sum = 0.0, norm;
        for (i = 0; i < n; i++) for (j = 0; j < n; j++) diff += fabs(x[i] - x[j]);
        for (i = 0; i < n; i++) sum += x[i];
        norm = 2 * n * sum;
        return (diff / (norm + 1e-10)) / ((n - 1) / (double)n);
    } else {
        double *sorted = (double *)alloca(sizeof(double) * n), tot = 0.0, cum = 0.0, acc = 0.0;
        memcpy(sorted, x, sizeof(double) * n);
        qsort(sorted, n, sizeof(double), cmp_double_idx);
        for (i = 0; i < n; i++) tot += sorted[i];
        for (i = 0; i < n; i++) { cum += sorted[i]; acc += cum / (tot + 1e-10); }
        return 1 - (2.0 / (n + 1)) * acc;
    }
}

static void current_metrics(const orc_batch *b, const env_t *s, double *out) { /* :269-318 */
    int A = b->cfg.n_agents, a;
    double *endow = (double *)alloca(sizeof(double) * A);
    double eta = b->cfg.isoelastic_eta, coef = energy_weight(b, s) * b->cfg.energy_cost;
    for (a = 0; a < A; a++) {
        double x = s->coin[a] + s->esc_coin[a];
        double util_c = (eta == 1.0) ? log(fmax(1.0, x)) : (pow(x, 1 - eta) - 1) / (1 - eta); /* rewards.py:35-40 */
        endow[a] = x;
        out[a] = util_c - s->labor[a] * coef;
    }
    if (b->cfg.planner_reward_type == ORC_SWF_COIN_EQ_TIMES_PROD) { /* rewards.py:84-101 */
        double sum = 0.0, eqw = 1 - b->cfg.mixing_weight_gini_vs_coin, prod, equality;
        for (a = 0; a < A; a++) sum += endow[a];
        prod = sum / A;
        equality = eqw * (1 - get_gini(endow, A)) + (1 - eqw);
        out[A] = equality * prod;
    } else { /* rewards.py:104-133 */
        double wsum = 0.0, acc = 0.0;
        for (a = 0; a < A; a++) wsum += 1 / fmax(endow[a], 1.0);
        for (a = 0; a < A; a++) {
            double w = (1 / fmax(endow[a], 1.0)) / wsum;
            acc += (b->cfg.planner_reward_type == ORC_SWF_INV_INCOME_COIN ? endow[a] : out[a]) * w;
        }
        out[A] = acc;
    }
}

static void compute_reward(const orc_batch *b, env_t *s) { /* :519-559 */
    int A = b->cfg.n_agents, a;
    double *prev = (double *)alloca(sizeof(double) * (A + 1)), avg = 0.0;
    memcpy(prev, s->curr_metric, sizeof(double) * (A + 1));
    current_metrics(b, s, s->curr_metric);
    for (a = 0; a <= A; a++) s->rew[a] = s->curr_metric[a] - prev[a];
    avg = np_sum(s->rew, A) / A; /* np.mean([...]) :552 - numpy's pairwise order: the sign of a sum that cancels to
                                  * rounding noise (a step of trades only) decides `avg > 0` */
    if (avg > 0) s->auto_warmup_integrator += 1;
}

/* ------------------------------------------------------------------------- */
/* observations + masks                                                       */
/* ------------------------------------------------------------------------- */

typedef struct { char key[96]; const double *v; int n; } field_t;
static int cmp_field(const void *x, const void *y) { return strcmp(((const field_t *)x)->key, ((const field_t *)y)->key); }

static const char *COMMODITY_NAME[2] = {"Stone", "Wood"};

/* BaseEnvironment._package: scalars/1-D fields concatenated in sorted key order -> float32 (base_env.py:562-612) */
static int pack_flat(field_t *f, int nf, float *out) {
    int i, j, n = 0;
    qsort(f, nf, sizeof(field_t), cmp_field);
    for (i = 0; i < nf; i++) for (j = 0; j < f[i].n; j++) out[n++] = (float)f[i].v[j];
    return n;
}

static void add_field(field_t *f, int *nf, const char *key, const double *v, int n) {
    snprintf(f[*nf].key, sizeof(f[*nf].key), "%s", key);
    f[*nf].v = v; f[*nf].n = n; (*nf)++;
}

static void generate_observations(const orc_batch *b, env_t *s) {
    const orc_config *cfg = &b->cfg;
    const orc_dims *d = &b->dims;
    int A = cfg->n_agents, H = cfg->height, W = cfg->width, HW = H * W, P = cfg->max_bid_ask + 1;
    int M = d->n_map_ch, win = d->win, w = cfg->obs_range;
    int a, c, k, p, ch, dr, dc, i;
    double inv_scale = cfg->allow_observation_scaling ? 0.01 : 1.0;
    double time_scale = cfg->allow_observation_scaling ? (double)cfg->episode_length : 1.0;
    double time_v = s->t / time_scale;
    const double *chan[6];
    int16_t *loc_map = (int16_t *)alloca(sizeof(int16_t) * HW);

    /* maps.state: stack in map-key order Stone, Wood, House, [Water], StoneSourceBlock, WoodSourceBlock (world.py:59-90, 314-317) */
    ch = 0;
    chan[ch++] = s->res[STONE]; chan[ch++] = s->res[WOOD]; chan[ch++] = s->house;
    if (cfg->has_water) chan[ch++] = s->water;
    chan[ch++] = s->src[STONE]; chan[ch++] = s->src[WOOD];

    /* world.loc_map, world.py:406-416 */
    for (k = 0; k < HW; k++) loc_map[k] = -1;
    for (a = 0; a < A; a++) loc_map[s->loc_r[a] * W + s->loc_c[a]] = (int16_t)a;

    /* planner spatial obs (layout_from_file.py:435-466): full map + idx maps (+2, 1 -> 0) */
    for (ch = 0; ch < M; ch++) for (k = 0; k < HW; k++) s->p_map[ch * HW + k] = (float)chan[ch][k];
    for (k = 0; k < HW; k++) {
        int16_t o = (int16_t)(s->owner[k] + 2), l = (int16_t)(loc_map[k] + 2);
        s->p_idx[k] = (o == 1) ? 0 : o;
        s->p_idx[HW + k] = (l == 1) ? 0 : l;
    }

    /* agent windows (layout_from_file.py:468-515): zero pad, extra channel = 1 inside padding?  No:
     * constant_values=[(0,1),...] pads ONE extra channel after the last with value 1 everywhere,
     * and the spatial padding of every channel (including that one) is 0. */
    if (cfg->full_observability) { /* layout_from_file.py:465-472: curr_map and a self-recoded copy of agent_idx_maps */
        for (a = 0; a < A; a++) {
            float *am = s->a_map + (size_t)a * d->a_map_elems;
            int16_t *ai = s->a_idx + (size_t)a * d->a_idx_elems;
            memcpy(am, s->p_map, sizeof(float) * M * HW);
            for (k = 0; k < 2 * HW; k++) ai[k] = (s->p_idx[k] == a + 2) ? 1 : s->p_idx[k];
        }
    } else
    for (a = 0; a < A; a++) {
        float *am = s->a_map + (size_t)a * (M + 1) * win * win;
        int16_t *ai = s->a_idx + (size_t)a * 2 * win * win;
        for (dr = 0; dr < win; dr++) for (dc = 0; dc < win; dc++) {
            int r = s->loc_r[a] + dr - w, cc = s->loc_c[a] + dc - w, o = dr * win + dc;
            int inside = (r >= 0 && r < H && cc >= 0 && cc < W);
            for (ch = 0; ch < M; ch++) am[ch * win * win + o] = inside ? (float)chan[ch][r * W + cc] : 0.0f;
            am[M * win * win + o] = inside ? 1.0f : 0.0f;
            {
                int16_t vo = inside ? s->p_idx[r * W + cc] : 0, vl = inside ? s->p_idx[HW + r * W + cc] : 0;
                if (vo == a + 2) vo = 1;
                if (vl == a + 2) vl = 1;
                ai[o] = vo; ai[win * win + o] = vl;
            }
        }
    }

    /* ---- scalar / vector fields ---- */
    {
        /* CDA obs (continuous_double_auction.py:491-542) */
        double net_hist[2][64], scaled_hist[2][64], market_rate[2], full_asks[2][64], full_bids[2][64];
        double avail_asks[64], avail_bids[64];
        double tax_is_tax_day = 0, tax_is_first_day = 0, tax_phase = 0;
        double zero3[3] = {0, 0, 0};
        if (b->has[ORC_COMP_CDA]) {
            for (c = 0; c < 2; c++) {
                double dot = 0.0, tot = 0.0;
                for (p = 0; p < P; p++) {
                    double acc = 0.0, fa = 0.0, fb = 0.0;
                    for (a = 0; a < A; a++) {
                        acc += s->price_hist[c][a * P + p];
                        fa += s->ask_hist[c][a * P + p];
                        fb += s->bid_hist[c][a * P + p];
                    }
                    net_hist[c][p] = acc; full_asks[c][p] = fa; full_bids[c][p] = fb;
                    scaled_hist[c][p] = acc * inv_scale;
                }
                for (p = 0; p < P; p++) { dot += p * net_hist[c][p]; tot += net_hist[c][p]; }
                market_rate[c] = dot / fmax(0.001, tot);
            }
        }
        if (b->has[ORC_COMP_TAX]) { /* redistribution.py:974-1023 */
            tax_is_tax_day = (s->tax_pos >= cfg->period) ? 1.0 : 0.0;
            tax_is_first_day = (s->tax_pos == 1) ? 1.0 : 0.0;
            tax_phase = (double)s->tax_pos / cfg->period;
        }

        for (a = 0; a < A; a++) {
            field_t f[48]; int nf = 0;
            field_t fp[16]; int nfp = 0;
            char key[96];
            double loc_row = (double)s->loc_r[a] / H, loc_col = (double)s->loc_c[a] / W;
            double inv_coin = s->coin[a] * inv_scale, inv_stone = s->inv[STONE][a] * inv_scale,
                   inv_wood = s->inv[WOOD][a] * inv_scale;
            double bp = 0, marg = 0;
            double my_av_asks[2][64], my_av_bids[2][64];
            if (!cfg->full_observability) {
                add_field(f, &nf, "world-loc-row", &loc_row, 1);
                add_field(f, &nf, "world-loc-col", &loc_col, 1);
            }
            add_field(f, &nf, "world-inventory-Coin", &inv_coin, 1);
            add_field(f, &nf, "world-inventory-Stone", &inv_stone, 1);
            add_field(f, &nf, "world-inventory-Wood", &inv_wood, 1);
            add_field(f, &nf, "time", &time_v, 1);
            if (!cfg->full_observability) { /* the p<i> entries are only set in the windowed branch (:509-513) */
                add_field(fp, &nfp, "world-inventory-Coin", &inv_coin, 1);
                add_field(fp, &nfp, "world-inventory-Stone", &inv_stone, 1);
                add_field(fp, &nfp, "world-inventory-Wood", &inv_wood, 1);
            }
            if (cfg->planner_gets_spatial_info && !cfg->full_observability) {
                add_field(fp, &nfp, "world-loc-row", &loc_row, 1);
                add_field(fp, &nfp, "world-loc-col", &loc_col, 1);
            }
            if (b->has[ORC_COMP_BUILD]) { /* build.py:163-178 */
                bp = s->build_payment[a] / cfg->build_payment;
                add_field(f, &nf, "Build-build_payment", &bp, 1);
                add_field(f, &nf, "Build-build_skill", &s->build_skill[a], 1);
            }
            if (b->has[ORC_COMP_GATHER]) add_field(f, &nf, "Gather-bonus_gather_prob", &s->bonus_prob[a], 1);
            if (b->has[ORC_COMP_CDA]) {
                for (c = 0; c < 2; c++) {
                    for (p = 0; p < P; p++) {
                        my_av_asks[c][p] = full_asks[c][p] - s->ask_hist[c][a * P + p];
                        my_av_bids[c][p] = full_bids[c][p] - s->bid_hist[c][a * P + p];
                    }
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-market_rate-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, &market_rate[c], 1);
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-price_history-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, scaled_hist[c], P);
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-available_asks-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, my_av_asks[c], P);
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-available_bids-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, my_av_bids[c], P);
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-my_asks-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, &s->ask_hist[c][a * P], P);
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-my_bids-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, &s->bid_hist[c][a * P], P);
                }
            }
            if (b->has[ORC_COMP_TAX]) {
                marg = marginal_rate(b, s, (s->coin[a] + s->esc_coin[a]) - s->last_coin[a]);
                add_field(f, &nf, "PeriodicBracketTax-is_tax_day", &tax_is_tax_day, 1);
                add_field(f, &nf, "PeriodicBracketTax-is_first_day", &tax_is_first_day, 1);
                add_field(f, &nf, "PeriodicBracketTax-tax_phase", &tax_phase, 1);
                add_field(f, &nf, "PeriodicBracketTax-last_incomes", s->last_income_obs_sorted, A);
                add_field(f, &nf, "PeriodicBracketTax-curr_rates", s->curr_rates_obs, cfg->n_brackets);
                add_field(f, &nf, "PeriodicBracketTax-marginal_rate", &marg, 1);
                add_field(fp, &nfp, "PeriodicBracketTax-last_income", &s->last_income_obs[a], 1);
                add_field(fp, &nfp, "PeriodicBracketTax-last_marginal_rate", &s->last_marg[a], 1);
                add_field(fp, &nfp, "PeriodicBracketTax-curr_marginal_rate", &marg, 1);
            }
            (void)avail_asks; (void)avail_bids;
            i = pack_flat(f, nf, s->a_flat + (size_t)a * d->flat_a);
            if (i != d->flat_a) { fprintf(stderr, "oracle: agent flat %d != %d\n", i, d->flat_a); abort(); }
            i = pack_flat(fp, nfp, s->p_agents + (size_t)a * d->flat_pa);
            if (i != d->flat_pa) { fprintf(stderr, "oracle: p<i> flat %d != %d\n", i, d->flat_pa); abort(); }
        }
        { /* planner */
            field_t f[32]; int nf = 0; char key[96];
            add_field(f, &nf, "world-inventory-Coin", &zero3[0], 1);
            add_field(f, &nf, "world-inventory-Stone", &zero3[1], 1);
            add_field(f, &nf, "world-inventory-Wood", &zero3[2], 1);
            add_field(f, &nf, "time", &time_v, 1);
            if (b->has[ORC_COMP_CDA]) {
                for (c = 0; c < 2; c++) {
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-market_rate-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, &market_rate[c], 1);
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-price_history-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, scaled_hist[c], P);
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-full_asks-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, full_asks[c], P);
                    snprintf(key, sizeof key, "ContinuousDoubleAuction-full_bids-%s", COMMODITY_NAME[c]);
                    add_field(f, &nf, key, full_bids[c], P);
                }
            }
            if (b->has[ORC_COMP_TAX]) {
                add_field(f, &nf, "PeriodicBracketTax-is_tax_day", &tax_is_tax_day, 1);
                add_field(f, &nf, "PeriodicBracketTax-is_first_day", &tax_is_first_day, 1);
                add_field(f, &nf, "PeriodicBracketTax-tax_phase", &tax_phase, 1);
                add_field(f, &nf, "PeriodicBracketTax-last_incomes", s->last_income_obs_sorted, A);
                add_field(f, &nf, "PeriodicBracketTax-curr_rates", s->curr_rates_obs, cfg->n_brackets);
            }
            i = pack_flat(f, nf, s->p_flat);
            if (i != d->flat_p) { fprintf(stderr, "oracle: planner flat %d != %d\n", i, d->flat_p); abort(); }
        }
    }
    s->time_obs = (float)time_v;

    /* ---- masks (base_env.py:706-756, base_agent.py:440-460) ---- */
    for (a = 0; a < A; a++) {
        float *m = s->a_mask + (size_t)a * d->mask_a;
        int n = 0, si;
        if (!cfg->multi_action_agents) m[n++] = 1.0f;
        for (si = 0; si < b->n_sub; si++) {
            if (cfg->multi_action_agents) m[n++] = 1.0f;
            if (b->sub_kind[si] == 0) { /* build.py:180-193 */
                m[n++] = agent_can_build(b, s, a) ? 1.0f : 0.0f;
            } else if (b->sub_kind[si] == 1) { /* Buy_c: continuous_double_auction.py:544-580 */
                c = b->sub_c[si];
                for (p = 0; p < P; p++)
                    m[n++] = (s->n_orders[c][a] < cfg->max_num_orders && (double)p <= s->coin[a]) ? 1.0f : 0.0f;
            } else if (b->sub_kind[si] == 2) { /* Sell_c */
                c = b->sub_c[si];
                for (p = 0; p < P; p++)
                    m[n++] = (s->n_orders[c][a] < cfg->max_num_orders && s->inv[c][a] > 0) ? 1.0f : 0.0f;
            } else { /* Gather mask move.py:167-188: [Left, Right, Up, Down], zero outside the world */
                static const int roff[4] = {0, 0, -1, 1}, coff[4] = {-1, 1, 0, 0};
                for (i = 0; i < 4; i++) {
                    int r = s->loc_r[a] + roff[i], cc = s->loc_c[a] + coff[i];
                    int ok = (r >= 0 && r < H && cc >= 0 && cc < W) && s->unoccupied[r * W + cc] && accessible(b, s, r, cc, a);
                    m[n++] = ok ? 1.0f : 0.0f;
                }
            }
        }
        if (n != d->mask_a) { fprintf(stderr, "oracle: mask %d != %d\n", n, d->mask_a); abort(); }
    }
    if (planner_has_tax_actions(cfg)) { /* redistribution.py:1025-1104 (multi-action planner) */
        int n = 0, R = cfg->n_disc_rates, bi, r;
        if (cfg->single_action_planner) s->p_mask[n++] = 1.0f; /* one global NO-OP (base_agent.py:452-453) */
        for (bi = 0; bi < cfg->n_brackets; bi++) {
            if (!cfg->single_action_planner) s->p_mask[n++] = 1.0f;
            for (r = 0; r < R; r++)
                s->p_mask[n++] = (s->tax_pos != 1) ? 0.0f : (float)s->planner_mask_rates[r];
        }
    } else {
        s->p_mask[0] = 1.0f;
    }
}

/* ------------------------------------------------------------------------- */
/* load / step                                                                */
/* ------------------------------------------------------------------------- */

int orc_load_env(orc_batch *b, int32_t e,
                 const uint8_t *stone, const uint8_t *wood, const uint8_t *stone_src,
                 const uint8_t *wood_src, const uint8_t *water, const int16_t *loc,
                 const double *coin, const int32_t *inv_stone, const int32_t *inv_wood,
                 const double *build_payment, const double *build_skill,
                 const double *bonus_gather_prob, const uint32_t *mt_key, int32_t mt_pos,
                 int32_t completions) {
    env_t *s;
    int A = b->cfg.n_agents, HW = b->cfg.height * b->cfg.width, P = b->cfg.max_bid_ask + 1, a, k, c, r;
    if (e < 0 || e >= b->n_envs) return -1;
    s = &b->envs[e];
    for (k = 0; k < HW; k++) {
        s->res[STONE][k] = stone[k]; s->res[WOOD][k] = wood[k];
        s->src[STONE][k] = stone_src[k]; s->src[WOOD][k] = wood_src[k];
        s->water[k] = water ? water[k] : 0;
        s->house[k] = 0; s->owner[k] = -1; s->unoccupied[k] = 1;
    }
    for (a = 0; a < A; a++) {
        s->loc_r[a] = loc[2 * a]; s->loc_c[a] = loc[2 * a + 1];
        s->unoccupied[s->loc_r[a] * b->cfg.width + s->loc_c[a]] = 0;
        s->coin[a] = coin[a]; s->esc_coin[a] = 0; s->labor[a] = 0;
        s->inv[STONE][a] = inv_stone ? inv_stone[a] : 0; s->inv[WOOD][a] = inv_wood ? inv_wood[a] : 0;
        s->esc[STONE][a] = 0; s->esc[WOOD][a] = 0;
        s->build_payment[a] = build_payment[a]; s->build_skill[a] = build_skill[a];
        s->bonus_prob[a] = bonus_gather_prob[a];
    }
    for (c = 0; c < 2; c++) { /* continuous_double_auction.py:643-668 */
        s->n_bids[c] = s->n_asks[c] = 0;
        for (a = 0; a < A; a++) s->n_orders[c][a] = 0;
        for (k = 0; k < A * P; k++) { s->price_hist[c][k] = 0; s->bid_hist[c][k] = 0; s->ask_hist[c][k] = 0; }
    }
    memcpy(s->mt, mt_key, sizeof(uint32_t) * 624);
    s->mt_pos = mt_pos;
    s->completions = completions;
    s->t = 0;
    s->done = 0;
    if (b->has[ORC_COMP_TAX]) tax_reset(b, s);
    /* component resets clear the episode logs (build.py:256, continuous_double_auction.py:664, redistribution.py:1131-1135) */
    memset(s->n_builds, 0, sizeof(double) * A);
    s->n_trades = 0;
    memset(s->trade_n, 0, sizeof(double) * A * 4); memset(s->trade_price, 0, sizeof(double) * A * 4);
    s->tax_periods = s->tax_collected = s->tax_eff_sum = 0;
    memset(s->tax_sched_sum, 0, sizeof(s->tax_sched_sum)); memset(s->tax_occupancy, 0, sizeof(s->tax_occupancy));
    memset(s->tax_income_pos, 0, sizeof(double) * A); memset(s->tax_paid, 0, sizeof(double) * A);
    /* planner "new_taxes" mask for this episode (redistribution.py:1051-1092) */
    for (r = 0; r < b->cfg.n_disc_rates; r++) {
        if (!b->cfg.tax_annealing) s->planner_mask_rates[r] = 1.0;
        else { /* components/utils.py:10-57, 60-115 */
            double full = 0.0, vis, lim;
            int q;
            for (q = 0; q < b->cfg.n_disc_rates; q++) full = fmax(full, fabs(b->cfg.disc_rates[q]));
            vis = fmax(0.0, fmin(1.0, b->cfg.annealing_slope * (completions - b->cfg.annealing_warmup)));
            lim = vis * full;
            s->planner_mask_rates[r] = (fabs(b->cfg.disc_rates[r]) <= lim) ? 1.0 : 0.0;
        }
    }
    /* scenario additional_reset_steps: metric_0 (layout_from_file.py:588-593) */
    current_metrics(b, s, s->curr_metric);
    for (a = 0; a <= A; a++) s->rew[a] = 0;
    generate_observations(b, s);
    return 0;
}

/* BaseAgent.parse_actions, base_agent.py:407-438 */
static void decode_actions(const orc_batch *b, env_t *s, const int32_t *act_a, const int32_t *act_p) {
    int A = b->cfg.n_agents, a, si, i;
    for (a = 0; a < A; a++) {
        s->act_build[a] = 0; s->act_move[a] = 0;
        s->act_buy[0][a] = s->act_buy[1][a] = s->act_sell[0][a] = s->act_sell[1][a] = 0;
        for (si = 0; si < b->n_sub; si++) {
            int v = 0;
            if (b->cfg.multi_action_agents) v = act_a ? act_a[a * b->n_sub + si] : 0;
            else { /* single_action_map: concatenated subspaces, 0 = global NO-OP */
                int g = act_a ? act_a[a] : 0, lo = 1, sj;
                for (sj = 0; sj < si; sj++) lo += b->sub_n[sj];
                if (g >= lo && g < lo + b->sub_n[si]) v = g - lo + 1;
            }
            if (b->sub_kind[si] == 0) s->act_build[a] = v;
            else if (b->sub_kind[si] == 1) s->act_buy[b->sub_c[si]][a] = v;
            else if (b->sub_kind[si] == 2) s->act_sell[b->sub_c[si]][a] = v;
            else s->act_move[a] = v;
        }
    }
    for (i = 0; i < ORC_MAX_BRACKETS; i++) s->act_tax[i] = 0;
    if (act_p && b->cfg.single_action_planner && b->dims.n_act_p == 1) {
        int g = act_p[0], R = b->cfg.n_disc_rates; /* single_action_map: bracket (g-1) / R, sub-action (g-1) % R + 1 */
        if (g >= 1 && g <= b->cfg.n_brackets * R) s->act_tax[(g - 1) / R] = (g - 1) % R + 1;
    } else if (act_p) for (i = 0; i < b->dims.n_act_p; i++) s->act_tax[i] = act_p[i];
}

static void step_env(const orc_batch *b, env_t *s, const int32_t *act_a, const int32_t *act_p) {
    int i;
    decode_actions(b, s, act_a, act_p);
    s->t += 1; /* base_env.py:1000 */
    for (i = 0; i < b->cfg.n_comp; i++) { /* :1002-1003 */
        switch (b->cfg.comp[i]) {
            case ORC_COMP_BUILD: build_step(b, s); break;
            case ORC_COMP_CDA: cda_step(b, s); break;
            case ORC_COMP_GATHER: gather_step(b, s); break;
            case ORC_COMP_TAX: tax_step(b, s); break;
            case ORC_COMP_WEALTH: wealth_step(b, s); break;
        }
    }
    scenario_step(b, s);          /* :1005 */
    generate_observations(b, s);  /* :1007-1010 */
    compute_reward(b, s);         /* :1011 */
    s->done = s->t >= b->cfg.episode_length; /* :1012 */
}

typedef struct { orc_batch *b; const int32_t *aa, *ap; int lo, hi, cpu; } job_t;
static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    if (j->cpu >= 0) {  /* cpu_baseline only: one thread per allowed core, each pinned to its own */
        cpu_set_t one;
        CPU_ZERO(&one); CPU_SET(j->cpu, &one);
        pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
    }
    int e, na = j->b->cfg.n_agents * j->b->dims.n_act_a, np = j->b->dims.n_act_p;
    for (e = j->lo; e < j->hi; e++)
        step_env(j->b, &j->b->envs[e], j->aa ? j->aa + (size_t)e * na : NULL,
                 (j->ap && np) ? j->ap + (size_t)e * np : NULL);
    return NULL;
}

int orc_step(orc_batch *b, const int32_t *actions_a, const int32_t *actions_p, int32_t n_threads) {
    int i;
    if (n_threads <= 1) {
        job_t j = {b, actions_a, actions_p, 0, b->n_envs, -1};
        worker(&j);
        return 0;
    }
    {
        pthread_t *th = (pthread_t *)alloca(sizeof(pthread_t) * n_threads);
        job_t *jobs = (job_t *)alloca(sizeof(job_t) * n_threads);
        cpu_set_t allowed;
        int n_allowed = 0, next_cpu = -1;
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) n_allowed = CPU_COUNT(&allowed);
        for (i = 0; i < n_threads; i++) {
            jobs[i].b = b; jobs[i].aa = actions_a; jobs[i].ap = actions_p;
            jobs[i].cpu = -1;
            if (n_allowed >= n_threads) {  /* the i-th allowed core */
                do { next_cpu++; } while (next_cpu < CPU_SETSIZE && !CPU_ISSET(next_cpu, &allowed));
                if (next_cpu < CPU_SETSIZE) jobs[i].cpu = next_cpu;
            }
            jobs[i].lo = (int)((long long)b->n_envs * i / n_threads);
            jobs[i].hi = (int)((long long)b->n_envs * (i + 1) / n_threads);
            pthread_create(&th[i], NULL, worker, &jobs[i]);
        }
        for (i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* getters                                                                    */
/* ------------------------------------------------------------------------- */

void orc_rng_words(orc_batch *b, int32_t e, uint32_t *out, int32_t n) {
    int i; for (i = 0; i < n; i++) out[i] = mt_next32(&b->envs[e]);
}
double orc_rng_rand(orc_batch *b, int32_t e) { return np_rand(&b->envs[e]); }
void orc_rng_permutation(orc_batch *b, int32_t e, int32_t n, int32_t *out) { np_permutation(&b->envs[e], n, out); }

int orc_get_obs(const orc_batch *b, int32_t e, float *a_map, int16_t *a_idx, float *a_flat, float *a_mask,
                float *p_map, int16_t *p_idx, float *p_flat, float *p_agents, float *p_mask,
                float *time_obs, double *rew, int32_t *done) {
    const env_t *s = &b->envs[e];
    const orc_dims *d = &b->dims;
    int A = b->cfg.n_agents, HW = b->cfg.height * b->cfg.width;
    if (a_map) memcpy(a_map, s->a_map, sizeof(float) * A * d->a_map_elems);
    if (a_idx) memcpy(a_idx, s->a_idx, sizeof(int16_t) * A * d->a_idx_elems);
    if (a_flat) memcpy(a_flat, s->a_flat, sizeof(float) * A * d->flat_a);
    if (a_mask) memcpy(a_mask, s->a_mask, sizeof(float) * A * d->mask_a);
    if (p_map) memcpy(p_map, s->p_map, sizeof(float) * d->n_map_ch * HW);
    if (p_idx) memcpy(p_idx, s->p_idx, sizeof(int16_t) * 2 * HW);
    if (p_flat) memcpy(p_flat, s->p_flat, sizeof(float) * d->flat_p);
    if (p_agents) memcpy(p_agents, s->p_agents, sizeof(float) * A * d->flat_pa);
    if (p_mask) memcpy(p_mask, s->p_mask, sizeof(float) * d->mask_p);
    if (time_obs) *time_obs = s->time_obs;
    if (rew) memcpy(rew, s->rew, sizeof(double) * (A + 1));
    if (done) *done = s->done;
    return 0;
}

/* Action masks of every env in one call (bench.py's host-side random policy; not part of the restated path). */
int orc_get_masks(const orc_batch *b, float *a_mask /* [E, A, mask_a] */, float *p_mask /* [E, mask_p] */) {
    const orc_dims *d = &b->dims;
    const size_t na = (size_t)b->cfg.n_agents * d->mask_a, np_ = (size_t)d->mask_p;
    int e;
    for (e = 0; e < b->n_envs; e++) {
        if (a_mask) memcpy(a_mask + (size_t)e * na, b->envs[e].a_mask, sizeof(float) * na);
        if (p_mask) memcpy(p_mask + (size_t)e * np_, b->envs[e].p_mask, sizeof(float) * np_);
    }
    return 0;
}

int orc_get_state(const orc_batch *b, int32_t e, uint8_t *cell, int8_t *owner, int16_t *loc,
                  double *coin, double *esc_coin, double *labor, int32_t *inv, int32_t *esc,
                  int32_t *n_orders, int32_t *bid_hist, int32_t *ask_hist, double *price_hist,
                  int32_t *tax_pos, int32_t *rate_idx, double *last_coin, double *last_income, double *last_marg,
                  uint32_t *mt_key, int32_t *mt_pos, int32_t *t) {
    const env_t *s = &b->envs[e];
    int A = b->cfg.n_agents, HW = b->cfg.height * b->cfg.width, P = b->cfg.max_bid_ask + 1, a, k, c, i;
    if (cell) for (k = 0; k < HW; k++)
        cell[k] = (uint8_t)((s->res[STONE][k] > 0) | ((s->res[WOOD][k] > 0) << 1) | ((s->src[STONE][k] > 0) << 2) |
                            ((s->src[WOOD][k] > 0) << 3) | ((s->water[k] > 0) << 4) | ((s->house[k] > 0) << 5));
    if (owner) for (k = 0; k < HW; k++) owner[k] = (int8_t)s->owner[k];
    for (a = 0; a < A; a++) {
        if (loc) { loc[2 * a] = (int16_t)s->loc_r[a]; loc[2 * a + 1] = (int16_t)s->loc_c[a]; }
        if (coin) coin[a] = s->coin[a];
        if (esc_coin) esc_coin[a] = s->esc_coin[a];
        if (labor) labor[a] = s->labor[a];
        if (inv) { inv[2 * a] = (int32_t)s->inv[STONE][a]; inv[2 * a + 1] = (int32_t)s->inv[WOOD][a]; }
        if (esc) { esc[2 * a] = (int32_t)s->esc[STONE][a]; esc[2 * a + 1] = (int32_t)s->esc[WOOD][a]; }
        if (last_coin) last_coin[a] = s->last_coin[a];
        if (last_income) last_income[a] = s->last_income[a];
        if (last_marg) last_marg[a] = s->last_marg[a];
    }
    for (c = 0; c < 2; c++) {
        for (a = 0; a < A; a++) if (n_orders) n_orders[c * A + a] = s->n_orders[c][a];
        for (k = 0; k < A * P; k++) {
            if (bid_hist) bid_hist[c * A * P + k] = (int32_t)s->bid_hist[c][k];
            if (ask_hist) ask_hist[c * A * P + k] = (int32_t)s->ask_hist[c][k];
            if (price_hist) price_hist[c * A * P + k] = s->price_hist[c][k];
        }
    }
    if (tax_pos) *tax_pos = s->tax_pos;
    if (rate_idx) for (i = 0; i < b->cfg.n_brackets; i++) rate_idx[i] = s->rate_idx[i];
    if (mt_key) memcpy(mt_key, s->mt, sizeof(uint32_t) * 624);
    if (mt_pos) *mt_pos = s->mt_pos;
    if (t) *t = s->t;
    return 0;
}

/* Episode statistics of env e, packed in the layout include/aie_b200.h documents for aie_dims.n_stats
 * (n = 1 + A + 8A [+ 35 + 2A with the tax component]); also the scenario's reward trackers. */
int orc_get_stats(const orc_batch *b, int32_t e, double *stats, double *curr_metric, int32_t *auto_warmup) {
    const env_t *s = &b->envs[e];
    int A = b->cfg.n_agents, a, i, n = 0;
    if (stats) {
        stats[n++] = s->n_trades;
        for (a = 0; a < A; a++) stats[n++] = s->n_builds[a];
        for (i = 0; i < 4 * A; i++) { stats[n++] = s->trade_n[i]; stats[n++] = s->trade_price[i]; }
        if (b->has[ORC_COMP_TAX]) {
            stats[n++] = s->tax_periods; stats[n++] = s->tax_collected; stats[n++] = s->tax_eff_sum;
            for (i = 0; i < 16; i++) stats[n++] = i < ORC_MAX_BRACKETS ? s->tax_sched_sum[i] : 0.0;
            for (i = 0; i < 16; i++) stats[n++] = i < ORC_MAX_BRACKETS ? s->tax_occupancy[i] : 0.0;
            for (a = 0; a < A; a++) stats[n++] = s->tax_income_pos[a];
            for (a = 0; a < A; a++) stats[n++] = s->tax_paid[a];
        }
    }
    if (curr_metric) memcpy(curr_metric, s->curr_metric, sizeof(double) * (A + 1));
    if (auto_warmup) *auto_warmup = s->auto_warmup_integrator;
    return n;
}

int orc_get_book(const orc_batch *b, int32_t e, int32_t c, int32_t side, int32_t *rows, int32_t cap) {
    const env_t *s = &b->envs[e];
    const order_t *v = side == 0 ? s->bids[c] : s->asks[c];
    int n = side == 0 ? s->n_bids[c] : s->n_asks[c], i;
    for (i = 0; i < n && i < cap; i++) { rows[3 * i] = v[i].agent; rows[3 * i + 1] = v[i].price; rows[3 * i + 2] = v[i].life; }
    return n;
}
