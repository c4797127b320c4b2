"""Oracle restatement of the fixed-step RK4 integrator (reference mlift/ode.py:9-53).

TEST INFRASTRUCTURE ONLY - see package docstring.
"""
import numpy as np
import torch


def compute_dt_seq(t_seq, dt):
    """Static step schedule (ode.py:45-53).

    Between consecutive output times take floor(gap/dt) full steps of `dt` plus, when
    the gap is not a multiple of dt, one final shorter step.  Returns the step sizes
    and, per output time, the index of the step that lands on it.
    """
    t_seq = np.asarray(t_seq, dtype=np.float64)
    gaps = t_seq[1:] - t_seq[:-1]
    n_full, rem = np.divmod(gaps, dt)
    n_full = n_full.astype(np.int64)
    has_rem = rem > 0
    n_full[has_rem] += 1
    last = n_full.cumsum() - 1
    dt_seq = np.full(last[-1] + 1, float(dt))
    dt_seq[last[has_rem]] = rem[has_rem]
    return dt_seq, last


def integrate_ode_rk4(dx_dt, x_init, t_seq, params, dt):
    """Classic RK4 over the static schedule; states at `t_seq` (ode.py:9-42)."""
    dt_seq, out_idx = compute_dt_seq(t_seq, float(np.asarray(dt).reshape(-1)[0]))
    x, t = x_init, float(t_seq[0])
    xs = []
    for h in dt_seq:
        h = float(h)
        k1 = dx_dt(x, t, params)
        k2 = dx_dt(x + h * k1 / 2, t + h / 2, params)
        k3 = dx_dt(x + h * k2 / 2, t + h / 2, params)
        k4 = dx_dt(x + h * k3, t + h, params)
        x = x + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6
        t = t + h
        xs.append(x)
    xs = torch.stack(xs)
    return torch.cat((x_init[None], xs[torch.as_tensor(out_idx)]))
