"""Host half of the Saez tax model (PeriodicBracketTax(tax_model="saez"), components/redistribution.py:437-823).

Device/host hybrid: the step kernel enacts taxes with whatever bracket rates sit in the "saez_rates"
field of an env's record, counts the (income, marginal rate) samples it has produced ("saez_n") and, while fewer than
500 exist, draws the reference's uniform random warm-up rates from the env's own numpy stream.  Once per tax period -
after the step that closed a period, i.e. right before the step whose tax_cycle_pos is 1 - this module appends the
period's samples to the replica's buffer and, when the buffer is full, evaluates the Saez formula (elasticity
regression, binned welfare weights / Pareto parameters, bracketisation) in float64 numpy exactly in the order the
reference does, then writes the new rates and their running average back into the record.

One estimator per replica, evaluated in a Python loop: exact semantics first (the reference runs the same numpy calls
once per env per period); at 8192 replicas this costs ~1 s per period.
"""
import numpy as np

BUFFER_SIZE = 500      # redistribution.py:276
N_BINS = 100           # :284


class SaezEstimator:
    def __init__(self, cutoffs, rate_min, rate_max, pareto_weight_type="inverse_income", fixed_elas=None):
        self.cutoffs = np.asarray(cutoffs, np.float64)
        self.B = len(self.cutoffs)
        self.bracket_sizes = np.concatenate([self.cutoffs, [np.inf]])[1:] - self.cutoffs   # :216-219
        self.rate_min, self.rate_max = float(rate_min), float(rate_max)
        assert pareto_weight_type in ("inverse_income", "uniform")
        self.pareto_weight_type = pareto_weight_type
        self.fixed_elas = None if fixed_elas is None else float(fixed_elas)
        self.edges = np.linspace(0, self.cutoffs[-1], N_BINS + 1)                          # :286-288
        self.bin_sizes = np.concatenate([self.edges[1:] - self.edges[:-1], [np.inf]])
        self.elas_tm1, self.elas_t, self.log_z0_tm1, self.log_z0_t = 0.5, 0.5, 0.0, 0.0    # :262-266
        self.running_avg = np.zeros(self.B)
        self.buffer = []                                                                    # [[income, marginal rate], ...]

    # -- the period's samples (:535-544) --------------------------------------------------------------
    def add_samples(self, incomes, marginal_rates):
        for z, tau in zip(incomes, marginal_rates):
            self.buffer.append([float(z), float(tau)])
        if len(self.buffer) > BUFFER_SIZE:
            del self.buffer[:len(self.buffer) - BUFFER_SIZE]

    @property
    def ready(self):
        return len(self.buffer) >= BUFFER_SIZE

    # -- elasticity (:549-599): OLS of log income on log(1 - rate), exponentially smoothed ------------------
    def _elasticity(self, data):
        keep = (data[:, 0] > 0) & (data[:, 1] < 1)
        zs, taus = data[keep, 0], data[keep, 1]
        if len(zs) < 10 or np.std(taus) < 1e-6:
            return float(self.elas_tm1), float(self.log_z0_tm1)
        x = np.log(np.maximum(1 - taus, 1e-9))
        X = np.stack([x, np.ones_like(x)]).T
        Y = np.log(np.maximum(zs, 1e-9))
        elas, log_z0 = np.linalg.inv(X.T.dot(X)).T.dot(X.T.dot(Y))
        return (1 - 0.98) * np.maximum(elas, 0.0) + 0.98 * self.elas_tm1, log_z0

    # -- binned welfare weights g(z) and Pareto parameters a(z) (:601-754) ---------------------------------
    def _pareto_weight(self, z):
        return np.ones_like(z) if self.pareto_weight_type == "uniform" else 1.0 / np.maximum(1, z)

    def _binned(self, incomes):
        counts, lefts = np.histogram(incomes, bins=self.edges)
        below, above = incomes[incomes < lefts[0]], incomes[incomes > lefts[-1]]
        n_total = np.sum(counts) + len(below) + len(above)
        pz = np.array([c / n_total for c in counts] + [len(above) / n_total])
        cum = [pz[0] + len(below) / n_total]
        for p in pz[1:]:
            cum.append(min(max(cum[-1] + p, 0), 1.0))
        cum = np.array(cum)
        # g(z): average normalised Pareto weight of everyone at or above z, bin-centred
        w_below = np.sum(self._pareto_weight(np.maximum(below, 0))) if len(below) else 0
        w_above = np.sum(self._pareto_weight(above)) if len(above) else 0
        w_bin = counts * self._pareto_weight(0.5 * (lefts[:-1] + lefts[1:]))
        norm = w_bin.sum() + w_below + w_above + 1e-9
        dens = np.concatenate([w_bin, [w_above]]) / norm
        g_geq = np.cumsum(dens[::-1])[::-1] / (np.cumsum(pz[::-1])[::-1] + 1e-9)
        gz = np.concatenate([0.5 * (g_geq[:-1] + g_geq[1:]), [g_geq[-1]]])
        # a(z) = z p(z) / P(Z >= z), per unit of bin width; the open top bin uses the mean income above the cutoff
        p_geq = 1 - cum + 0.5 * pz
        az = []
        for i in range(N_BINS):
            if pz[i] == 0:
                az.append(np.nan)
            else:
                z = 0.5 * (lefts[i] + lefts[i + 1])
                az.append(z * pz[i] / (min(max(p_geq[i], 0), 1) + 1e-9) / (lefts[i + 1] - lefts[i]))
        if len(above):
            m = np.mean(above)
            az.append(m / (m - lefts[-1] + 1e-9))
        else:
            az.append(0.0)
        return gz, np.array(az)

    # -- marginal rates per bin (:756-785), empty bins linearly interpolated ------------------------------
    @staticmethod
    def _bin_rates(gz, az, elas):
        taus = (1.0 - gz) / (1.0 - gz + az * elas + 1e-9)
        last_rate, last_i = 0.0, -1
        for i, tau in enumerate(taus):
            if np.isnan(tau):
                continue
            if i - last_i > 1:
                assert i != 0
                gap = list(range(last_i + 1, i))
                for k, r in zip(gap, np.linspace(last_rate, tau, len(gap) + 2)[1:-1]):
                    taus[k] = r
            last_rate, last_i = float(tau), i
        return taus

    # -- average marginal rate inside each tax bracket (:787-823) ------------------------------------------
    def _bracketize(self, bin_rates):
        out, last_total = [], 0
        for b, income in enumerate(self.cutoffs[1:]):
            due = np.maximum(0, np.sum(bin_rates * np.minimum(self.bin_sizes, np.maximum(0, income - self.edges))))
            out.append((due - last_total) / self.bracket_sizes[b])
            last_total = due
        out.append(bin_rates[-1])
        return np.array(out)

    def new_period_rates(self):
        """compute_and_set_new_period_rates_from_saez_formula (:437-511) for a full buffer -> bracket rates [B]."""
        data = np.array(self.buffer)
        self.elas_tm1, self.log_z0_tm1 = float(self.elas_t), float(self.log_z0_t)
        elas_t, log_z0_t = self._elasticity(data)
        self.elas_t, self.log_z0_t = float(elas_t), float(log_z0_t)
        if self.fixed_elas is not None:
            elas_t = self.fixed_elas
        gz, az = self._binned(data[:, 0])
        rates = np.clip(self._bracketize(self._bin_rates(gz, az, elas_t)), self.rate_min, self.rate_max)
        self.running_avg = self.running_avg * 0.99 + rates * 0.01
        return rates


class SaezHost:
    """The per-replica estimators of a BatchedFoundationEnv plus the traffic to / from the state records."""

    def __init__(self, env, tax_component):
        t = tax_component
        self.env = env
        self.A = env.n_agents
        self.est = [SaezEstimator(t.bracket_cutoffs, t.rate_min, t.rate_max, t.pareto_weight_type, t.saez_fixed_elas)
                    for _ in range(env.n_envs)]
        self._seen = np.zeros(env.n_envs, np.int64)   # samples already copied from the device, per replica
        self.elas_at_episode_end = [None] * env.n_envs  # saez/estimated_elasticity of previous_episode_metrics

    def _np(self, name, final=False):
        st = self.env.stepper
        return st.to_numpy(st.state_view(name, final=final))

    def _write_rows(self, name, rows, values):
        """state_view(name)[rows] = values (numpy or torch backed)."""
        v = self.env.stepper.state_view(name)
        if isinstance(v, np.ndarray):
            v[rows] = values
        else:
            import torch
            v[torch.as_tensor(rows, device=v.device)] = torch.as_tensor(values, device=v.device, dtype=v.dtype)

    def before_host_reset(self):
        return self._np("saez_n").copy(), self._np("saez_rates").copy()

    def after_host_reset(self, saved):
        """An explicit env.reset() repacks the records: put the persistent Saez state back.  As in the reference
        (redistribution.py:1123 vs :1138-1139) the reset observation's curr_rates are the previous episode's last rates,
        while the rates in force become the running average."""
        saez_n, old_rates = saved
        E = self.env.n_envs
        avg = np.zeros((E, 16))
        for e in range(E):
            avg[e, :self.est[e].B] = self.est[e].running_avg
        self._write_rows("saez_n", np.arange(E), saez_n.astype(np.int32))
        self._write_rows("saez_obs_rates", np.arange(E), old_rates)
        self._write_rows("saez_avg_rates", np.arange(E), avg)
        self._write_rows("saez_rates", np.arange(E), avg)
        if old_rates.any():
            self.env.stepper.observe()
        self.after_step()

    def after_step(self):
        """Call after every step (and after a reset): pulls new samples and, for replicas about to start a tax
        period with a full buffer, writes that period's rates."""
        n_dev = self._np("saez_n").astype(np.int64)
        fresh = np.nonzero(n_dev > self._seen)[0]
        if len(fresh):
            inc, marg = self._np("last_income"), self._np("last_marg")
            # a replica whose episode ended on that tax day has already been reset on the device (trackers zeroed):
            # its samples are in the end-of-episode snapshot
            ended = self._np("t")[fresh] == 0
            if ended.any():
                inc_f, marg_f = self._np("last_income", final=True), self._np("last_marg", final=True)
            for e, was_reset in zip(fresh, ended):  # one tax day per step at most: the A samples of the day that just closed
                if was_reset:
                    self.est[e].add_samples(inc_f[e], marg_f[e])
                else:
                    self.est[e].add_samples(inc[e], marg[e])
            self._seen[fresh] = n_dev[fresh]
        for e in np.nonzero(self._np("t") == 0)[0]:   # replicas that just finished an episode: metrics are taken before
            self.elas_at_episode_end[e] = float(self.est[e].elas_tm1)  # the new episode's first estimate
        # replicas about to run a tax_cycle_pos == 1 step (a finished episode without auto-reset waits for its reset)
        starting = np.nonzero((self._np("tax_pos") == 1) & (n_dev >= BUFFER_SIZE) &
                              (self._np("t") < self.env.episode_length))[0]
        if len(starting) == 0:
            return
        rates, avg = np.zeros((len(starting), 16)), np.zeros((len(starting), 16))
        for i, e in enumerate(starting):
            r = self.est[e].new_period_rates()
            rates[i, :len(r)] = r
            avg[i, :len(r)] = self.est[e].running_avg
        self._write_rows("saez_rates", starting, rates)
        self._write_rows("saez_avg_rates", starting, avg)
