"""ORACLE (test infrastructure only): CPU restatement of the reference's sampler loops around an arbitrary
eps-model callable.  PLMS: ldm/models/diffusion/plms.py:118-240; DDIM: ldm/models/diffusion/ddim.py:117-220;
generalized_steps: ddim/functions/denoising.py:10-32.  Schedules: util.py:21-74, ddpm.py:118-146."""
import numpy as np
import torch


def ldm_schedule(n_timestep=1000, linear_start=1e-4, linear_end=2e-2):
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()
    ac = np.cumprod(1.0 - betas, axis=0)
    return torch.tensor(ac, dtype=torch.float32)


def ddim_params(alphas_cumprod, S, eta=0.0):
    T = alphas_cumprod.shape[0]
    c = T // S
    steps = np.asarray(list(range(0, T, c))) + 1
    ac = alphas_cumprod.cpu()
    alphas = ac[steps]
    alphas_prev = np.asarray([ac[0]] + ac[steps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return steps, alphas, alphas_prev, sigmas, np.sqrt(1. - alphas)


def ddim_sample(model, x_T, cond, uc, scale, alphas_cumprod, S, eta=0.0, noises=None):
    """DDIMSampler.sample -> ddim_sampling -> p_sample_ddim (ldm/models/diffusion/ddim.py:57-220), temperature 1,
    no mask / score corrector.  noises[i] = the torch.randn draw of step i (ddim.py:217 via util.py:264-267);
    the reference draws it on every step, also when sigma is 0."""
    steps, alphas, alphas_prev, sigmas, sqrt_1ma = ddim_params(alphas_cumprod, S, eta)
    b = x_T.shape[0]
    img = x_T.clone()
    time_range = np.flip(steps)
    total = steps.shape[0]
    for i, step in enumerate(time_range):
        index = total - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        if uc is None or scale == 1.0:
            e_t = model(img, ts, cond)
        else:
            e_u, e_c = model(torch.cat([img] * 2), torch.cat([ts] * 2), torch.cat([uc, cond])).chunk(2)
            e_t = e_u + scale * (e_c - e_u)
        a_t = torch.full((b, 1, 1, 1), float(alphas[index]))
        a_prev = torch.full((b, 1, 1, 1), float(alphas_prev[index]))
        sigma_t = torch.full((b, 1, 1, 1), float(sigmas[index]))
        s1m = torch.full((b, 1, 1, 1), float(sqrt_1ma[index]))
        pred_x0 = (img - s1m * e_t) / a_t.sqrt()
        dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
        noise = sigma_t * (noises[i] if noises is not None else torch.zeros_like(img))
        img = a_prev.sqrt() * pred_x0 + dir_xt + noise
    return img


def plms_sample(model, x_T, cond, uc, scale, alphas_cumprod, S):
    """model(x, t, context) -> eps.  Mirrors plms_sampling + p_sample_plms (eta = 0)."""
    steps, alphas, alphas_prev, sigmas, sqrt_1ma = ddim_params(alphas_cumprod, S)
    b = x_T.shape[0]
    img = x_T.clone()
    time_range = np.flip(steps)
    total = steps.shape[0]
    old_eps = []

    def eps_of(x, t):
        if uc is None or scale == 1.0:
            return model(x, t, cond)
        e_u, e_c = model(torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([uc, cond])).chunk(2)
        return e_u + scale * (e_c - e_u)

    def x_prev_of(x, e, index):
        a_t, a_prev = float(alphas[index]), float(alphas_prev[index])
        pred_x0 = (x - float(sqrt_1ma[index]) * e) / np.sqrt(a_t)
        return float(np.sqrt(a_prev)) * pred_x0 + float(np.sqrt(1. - a_prev)) * e

    for i, step in enumerate(time_range):
        index = total - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        ts_next = torch.full((b,), int(time_range[min(i + 1, len(time_range) - 1)]), dtype=torch.long)
        e_t = eps_of(img, ts)
        if len(old_eps) == 0:
            x_prev = x_prev_of(img, e_t, index)
            e_t_next = eps_of(x_prev, ts_next)
            e_p = (e_t + e_t_next) / 2
        elif len(old_eps) == 1:
            e_p = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_p = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_p = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        img = x_prev_of(img, e_p, index)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
    return img


def generalized_steps(model, x, seq, betas, eta=0.0, noises=None):
    """ddim/functions/denoising.py:10-32; returns the final x."""
    n = x.size(0)
    beta = torch.cat([torch.zeros(1), betas.float()], dim=0)
    acp = (1 - beta).cumprod(dim=0)
    seq_next = [-1] + list(seq[:-1])
    xt = x
    for k, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
        t = torch.ones(n) * i
        at, at_next = acp[int(i) + 1], acp[int(j) + 1]
        et = model(xt, t)
        x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
        c1 = eta * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
        c2 = ((1 - at_next) - c1 ** 2).sqrt()
        noise = noises[k] if noises is not None else torch.zeros_like(x)
        xt = at_next.sqrt() * x0_t + c1 * noise + c2 * et
    return xt


# ------------------------------------------------------------------------------------------- DPM-Solver++ (2M)
class _DiscreteVP:
    """NoiseScheduleVP('discrete', alphas_cumprod=...) (ldm/models/diffusion/dpm_solver/dpm_solver.py:96-108,125-156):
    log(alpha_t) is piecewise linear in t over t_n = n / N, n = 1..N; fp32 like the reference."""

    def __init__(self, alphas_cumprod):
        self.log_alpha = 0.5 * torch.log(alphas_cumprod)
        self.total_N = self.log_alpha.shape[0]
        self.t_array = torch.linspace(0., 1., self.total_N + 1, dtype=alphas_cumprod.dtype)[1:]

    def marginal_log_mean_coeff(self, t):
        """interpolate_fn (:1132-1170): linear interpolation, the outermost segments extended beyond the key points."""
        K = self.total_N
        idx = torch.searchsorted(self.t_array, t.reshape(-1)).clamp(1, K - 1)
        x0, x1 = self.t_array[idx - 1], self.t_array[idx]
        y0, y1 = self.log_alpha[idx - 1], self.log_alpha[idx]
        return (y0 + (t.reshape(-1) - x0) * (y1 - y0) / (x1 - x0)).reshape(t.shape)

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * torch.log(1. - torch.exp(2. * lm))


def dpm_solver_sample(model, x_T, cond, uc, scale, alphas_cumprod, S):
    """DPMSolverSampler.sample (ldm/models/diffusion/dpm_solver/sampler.py:24-82): DPM-Solver++ with data prediction
    (predict_x0=True), multistep order 2, uniform time steps, lower_order_final, classifier-free guidance
    (dpm_solver.py:321-345 model_fn, :386-399 data prediction, :504-527 first-order update, :755-795 second-order update,
    :1077-1105 the multistep loop).  model(x, t_model, context) -> eps, t_model = (t - 1/N) * 1000 (:278-287)."""
    ns = _DiscreteVP(alphas_cumprod)
    dt = x_T.dtype
    b = x_T.shape[0]
    t_0, t_T = 1. / ns.total_N, 1.
    timesteps = torch.linspace(t_T, t_0, S + 1, dtype=dt)

    def data_pred(x, t):
        vec_t = t.expand(b)
        t_in = (vec_t - 1. / ns.total_N) * 1000.
        if uc is None or scale == 1.0:
            noise = model(x, t_in, cond)
        else:
            e_u, e_c = model(torch.cat([x] * 2), torch.cat([t_in] * 2), torch.cat([uc, cond])).chunk(2)
            noise = e_u + scale * (e_c - e_u)
        a, sg = ns.marginal_alpha(vec_t), ns.marginal_std(vec_t)
        return (x - sg.reshape(-1, 1, 1, 1) * noise) / a.reshape(-1, 1, 1, 1)

    def first(x, s, t, m_s):
        h = ns.marginal_lambda(t) - ns.marginal_lambda(s)
        return (ns.marginal_std(t) / ns.marginal_std(s)) * x - (ns.marginal_alpha(t) * torch.expm1(-h)) * m_s

    def second(x, m1, m0, t1, t0, t):
        l1, l0, lt = ns.marginal_lambda(t1), ns.marginal_lambda(t0), ns.marginal_lambda(t)
        h_0, h = l0 - l1, lt - l0
        r0 = h_0 / h
        D1 = (1. / r0) * (m0 - m1)
        k = ns.marginal_alpha(t) * (torch.exp(-h) - 1.)
        return (ns.marginal_std(t) / ns.marginal_std(t0)) * x - k * m0 - 0.5 * k * D1

    x = x_T.clone()
    ms, ts = [data_pred(x, timesteps[0])], [timesteps[0]]
    x = first(x, ts[-1], timesteps[1], ms[-1])                # init_order = 1
    ms.append(data_pred(x, timesteps[1]))
    ts.append(timesteps[1])
    for step in range(2, S + 1):
        t = timesteps[step]
        order = min(2, S + 1 - step) if S < 15 else 2
        x = first(x, ts[-1], t, ms[-1]) if order == 1 else second(x, ms[0], ms[1], ts[0], ts[1], t)
        ms[0], ts[0] = ms[1], ts[1]
        ts[1] = t
        if step < S:
            ms[1] = data_pred(x, t)
    return x
