// aie_covid_abi.inl — extern "C" entry points of the COVID-19 scenario (include/aie_b200.h, second half).
// Included after aie_abi.inl by both backends, which provide in namespace aie::be:
//   void *const_upload(const void *host, size_t bytes);  void const_free(void *dev);
//   int covid_launch_reset(aie_covid_env *, void *stream);  int covid_launch_step(aie_covid_env *, void *stream);
//   int covid_launch_sample(aie_covid_env *, uint64_t key, void *stream);
struct aie_covid_env {
    aie::CovidCfg cfg;
    aie::CovidBufs bufs;
    int n_envs, device;
    bool bound, loaded;
    int64_t launches;
    uint64_t sample_calls;
    std::vector<void *> owned;
};

extern "C" {

int aie_covid_create(const aie_covid_config *u, int32_t n_envs, int32_t device, aie_covid_env **out) {
    if (!u || !out) return fail(AIE_EINVAL, "null argument");
    if (u->abi_version != AIE_ABI_VERSION) return fail(AIE_EINVAL, "aie_covid_create: abi_version mismatch");
    if (n_envs < 1) return fail(AIE_EINVAL, "aie_covid_create: n_envs must be >= 1");
    if (u->n_states < 1 || u->n_states > 64) return fail(AIE_EINVAL, "aie_covid_create: n_states must be in [1, 64]");
    if (u->num_filters < 1 || u->num_filters > 8) return fail(AIE_EINVAL, "aie_covid_create: num_filters must be in [1, 8]");
    if (u->filter_len < 1 || u->num_filters < 1 || u->beta_delay < 1 || u->beta_delay > u->filter_len)
        return fail(AIE_EINVAL, "aie_covid_create: bad filter / delay sizes");
    if (u->episode_length < 1 || u->subsidy_interval < 1 || u->delivery_interval < 1 || u->num_subsidy_levels < 1 ||
        u->num_stringency_levels < 2 || u->num_stringency_levels > 100)
        return fail(AIE_EINVAL, "aie_covid_create: bad schedule parameters");
    if (u->start_date_index < 0 || u->start_date_index >= u->rw_policy_days)
        return fail(AIE_EINVAL, "aie_covid_create: start_date_index outside the real-world policy series");
    const void *req[] = {u->population, u->num_vaccines_per_delivery, u->beta_slopes, u->beta_intercepts, u->unemployment_bias,
                         u->daily_production_per_worker, u->maximum_productivity, u->agents_health_norm,
                         u->agents_economic_norm, u->min_agent_health, u->max_agent_health, u->min_agent_econ,
                         u->max_agent_econ, u->w_agent_health, u->w_agent_econ, u->conv_weights, u->conv_filters,
                         u->max_daily_subsidy_per_state, u->rw_policy, u->init_state};
    for (const void *p : req) if (!p) return fail(AIE_EINVAL, "aie_covid_create: a parameter array is NULL");
    {
        const int rc_dev = aie::be::check_device(device);   // same rules as aie_create: a real sm_100 device, no CPU path
        if (rc_dev != AIE_OK) return rc_dev;
    }
    AIE_DEVICE_SCOPE(device);
    aie_covid_env *env = new (std::nothrow) aie_covid_env();
    if (!env) return fail(AIE_ENOMEM, "out of host memory");
    env->n_envs = n_envs; env->device = device; env->bound = env->loaded = false; env->launches = 0; env->sample_calls = 0;
    memset(&env->bufs, 0, sizeof(env->bufs));
    aie::CovidCfg &c = env->cfg;
    memset(&c, 0, sizeof(c));
    const int S = u->n_states;
    c.S = S; c.T = u->episode_length; c.levels = u->num_stringency_levels; c.cooldown = u->action_cooldown_period;
    c.sub_interval = u->subsidy_interval; c.sub_levels = u->num_subsidy_levels;
    c.vac_begin = u->time_when_vaccine_delivery_begins; c.vac_interval = u->delivery_interval;
    c.t_first_delivery = u->t_first_delivery; c.beta_delay = u->beta_delay; c.L = u->filter_len; c.F = u->num_filters;
    c.sdi = u->start_date_index; c.rw_days = u->rw_policy_days; c.value_of_life = u->value_of_life;
    c.auto_reset = u->auto_reset ? 1 : 0; c.n_envs = n_envs;
    c.gamma = u->gamma; c.death_rate = u->death_rate; c.sick_rate = u->infection_too_sick_to_work_rate;
    c.p1865 = u->pop_between_age_18_65; c.rfr = u->risk_free_interest_rate; c.crra_eta = u->crra_eta;
    c.planner_health_norm = u->planner_health_norm; c.planner_econ_norm = u->planner_economic_norm;
    c.min_ph = u->min_planner_health; c.max_ph = u->max_planner_health; c.min_pe = u->min_planner_econ;
    c.max_pe = u->max_planner_econ; c.w_ph = u->w_planner_health; c.w_pe = u->w_planner_econ;
    c.reward_norm = u->reward_normalization_factor; c.time_scale = u->time_scale;
    c.dppw = u->daily_production_per_worker[0];
    bool ok = true;
    auto up = [&](const void *h, size_t bytes) -> const void * {
        void *d = aie::be::const_upload(h, bytes);
        if (!d) ok = false; else env->owned.push_back(d);
        return d;
    };
    c.pop = (const int32_t *)up(u->population, 4 * S);
    c.vac_per_delivery = (const int32_t *)up(u->num_vaccines_per_delivery, 4 * S);
    c.beta_slopes = (const float *)up(u->beta_slopes, 4 * S);
    c.beta_intercepts = (const float *)up(u->beta_intercepts, 4 * S);
    c.unemp_bias = (const float *)up(u->unemployment_bias, 4 * S);
    c.max_prod = (const float *)up(u->maximum_productivity, 4 * S);
    c.health_norm = (const float *)up(u->agents_health_norm, 4 * S);
    c.econ_norm = (const float *)up(u->agents_economic_norm, 4 * S);
    c.min_ah = (const float *)up(u->min_agent_health, 4 * S); c.max_ah = (const float *)up(u->max_agent_health, 4 * S);
    c.min_ae = (const float *)up(u->min_agent_econ, 4 * S); c.max_ae = (const float *)up(u->max_agent_econ, 4 * S);
    c.w_ah = (const float *)up(u->w_agent_health, 4 * S); c.w_ae = (const float *)up(u->w_agent_econ, 4 * S);
    c.conv_w = (const float *)up(u->conv_weights, 4 * (size_t)S * c.F);
    c.conv_filt = (const float *)up(u->conv_filters, 4 * (size_t)c.F * c.L);
    c.init_state = (const float *)up(u->init_state, 4 * 6 * (size_t)S);
    c.max_daily_subsidy = (const double *)up(u->max_daily_subsidy_per_state, 8 * S);
    c.rw_policy = (const int8_t *)up(u->rw_policy, (size_t)c.rw_days * S);
    if (!ok) { aie_covid_destroy(env); return fail(AIE_ECUDA, "aie_covid_create: uploading the parameters failed (no CUDA device? there is no CPU fallback)"); }
    *out = env;
    return AIE_OK;
}

int aie_covid_destroy(aie_covid_env *env) {
    if (!env) return AIE_OK;
    aie::be::DevScope dev_scope_(env->device);
    for (void *d : env->owned) aie::be::const_free(d);
    delete env;
    return AIE_OK;
}

int aie_covid_bind_buffers(aie_covid_env *env, const aie_covid_buffers *b) {
    if (!env || !b) return fail(AIE_EINVAL, "null argument");
    if (!b->state || !b->ints || !b->hdr || !b->ring || !b->actions_agent || !b->actions_planner || !b->obs_agent_state ||
        !b->obs_postsubsidy || !b->obs_lagged_stringency || !b->obs_policy_indicators || !b->obs_scalars || !b->mask_agent ||
        !b->mask_planner || !b->reward_agent || !b->reward_planner || !b->done)
        return fail(AIE_EINVAL, "aie_covid_bind_buffers: a required buffer is NULL");
    aie::CovidBufs &d = env->bufs;
    d.state = b->state; d.ints = b->ints; d.hdr = b->hdr; d.ring = b->ring; d.act_a = b->actions_agent; d.act_p = b->actions_planner;
    d.o_state = b->obs_agent_state; d.o_post = b->obs_postsubsidy; d.o_lag = b->obs_lagged_stringency;
    d.o_pol = b->obs_policy_indicators; d.o_scal = b->obs_scalars; d.mask_a = b->mask_agent; d.mask_p = b->mask_planner;
    d.rew_a = b->reward_agent; d.rew_p = b->reward_planner; d.done = b->done;
    d.chg = b->changes;   // optional
    env->bound = true;
    return AIE_OK;
}

int aie_covid_reset(aie_covid_env *env, void *stream) {
    if (!env) return fail(AIE_EINVAL, "null argument");
    if (!env->bound) return fail(AIE_ESTATE, "aie_covid_reset: buffers not bound");
    AIE_DEVICE_SCOPE(env->device);
    int rc = aie::be::covid_launch_reset(env, stream);
    if (rc == AIE_OK) env->loaded = true;
    return rc;
}

int aie_covid_step(aie_covid_env *env, void *stream) {
    if (!env) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->loaded) return fail(AIE_ESTATE, "aie_covid_step: bind buffers and reset first");
    AIE_DEVICE_SCOPE(env->device);
    return aie::be::covid_launch_step(env, stream);
}

int aie_covid_sample_random_actions(aie_covid_env *env, uint64_t seed, void *stream) {
    if (!env) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->loaded) return fail(AIE_ESTATE, "aie_covid_sample_random_actions: bind buffers and reset first");
    AIE_DEVICE_SCOPE(env->device);
    return aie::be::covid_launch_sample(env, aie::host_mix64(seed) ^ aie::host_mix64(++env->sample_calls), stream);
}

int64_t aie_covid_launch_count(const aie_covid_env *env) { return env ? env->launches : 0; }

}  // extern "C"
