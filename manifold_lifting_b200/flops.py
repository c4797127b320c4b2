"""Analytic algorithmic FP64 operation counts of the constrained leapfrog step (the single
source for bench.py's FP64 roofline; DESIGN.md section "Kernels and rooflines").

Counting rule: add / sub / mul / div / sqrt / exp / log / abs-max-compare = 1, fma = 2.
Counts follow the reference's formulas (systems.py:712-794 for the hierarchical Woodbury
family, systems.py:582-609 for the dense additive family), NOT the instruction stream
(ncu's executed-instruction counters are the cross-check, see profiles/).
"""


def hier_counts(U, Y):
    Q = U + 2 * Y
    c = {}
    c["constr"] = 5 * Y + 1
    c["jacob"] = c["constr"] + Y
    c["decompose"] = 3 * Y + U * Y + 2 * U * U * Y + U + (U * U * U) // 3 + 2 * U
    c["inv_gram"] = Y + 2 * U * Y + 2 * U * U + 2 * U * Y + 2 * Y
    c["rmult"] = 2 * U * Y + 2 * Y
    c["lmult"] = 2 * U * Y + 4 * Y
    c["pinv"] = c["inv_gram"] + c["rmult"]
    c["project"] = c["lmult"] + c["pinv"] + Q
    c["inv_jacprod"] = (3 * Y + U * Y + 2 * U * U * Y + U + Y + 2 * U * Y
                        + (2 * U * U * U) // 3 + 2 * U * U + 2 * U * Y + 2 * Y)  # fmt: skip
    c["newton_iter"] = c["jacob"] + Y + c["inv_jacprod"] + c["rmult"] + 2 * Q + Q
    c["quasi_newton_iter"] = c["constr"] + Y + c["pinv"] + 2 * Q + Q
    c["logdet"] = 2 * (U + Y) + 2
    c["nld"] = 4 * Y + 12
    c["grad_logdet"] = Y * (2 * U * U + 2 * U + 3 * U + 8) + 10
    c["evaluate"] = c["jacob"] + c["decompose"] + c["logdet"] + c["nld"] + c["grad_logdet"]
    c["step_fixed"] = (2 * Q + c["project"] + 2 * Q + 2 * Q + c["evaluate"] + c["project"]
                       + 2 * Q + 2 * Q + 2 * Q + c["project"])  # fmt: skip
    return c


def toy_counts():
    c = {"constr": 9, "jacob": 16, "decompose": 7, "inv_gram": 2, "rmult": 3, "lmult": 6}
    c["pinv"] = c["inv_gram"] + c["rmult"]
    c["project"] = c["lmult"] + c["pinv"] + 3
    c["inv_jacprod"] = 8
    c["newton_iter"] = c["jacob"] + 1 + c["inv_jacprod"] + c["rmult"] + 9
    c["quasi_newton_iter"] = c["constr"] + 1 + c["pinv"] + 9
    c["logdet"], c["nld"], c["grad_logdet"] = 1, 6, 12
    c["evaluate"] = c["jacob"] + c["decompose"] + c["logdet"] + c["nld"] + c["grad_logdet"]
    c["step_fixed"] = 6 + c["project"] + 6 + 6 + c["evaluate"] + c["project"] + 18 + c["project"]
    return c


def fn_counts(Y, n_steps, U=9, ND=8):
    """FitzHugh-Nagumo (streaming family).  Per RK4 stage: primal RHS 12 + RK4 update 8;
    per tangent direction J_f s (8) + forcing (2) + update (8); per second-order pair
    J_f X (8) + forcing (8) + update (8).  Second-order work is counted ONCE per pair
    (the kernel recomputes the first-order part once per pair group; that redundancy is
    not algorithmic work)."""
    NP = ND * (ND + 1) // 2
    stage0, stage1, stage2 = 20, 3 + 18 * ND, 24 * NP
    c = {}
    c["constr"] = 4 * n_steps * stage0 + 3 * Y
    c["jacob"] = 4 * n_steps * (stage0 + stage1) + 5 * Y
    c["second_order"] = 4 * n_steps * (stage0 + stage1 + stage2) + 4 * NP * Y
    c["decompose"] = Y * (U * (U + 1) + U + 8) + U * U * U // 3 + 20 * U
    c["inv_gram"] = Y * (2 * U + 8) + 2 * U * U + Y * (2 * U + 8)
    c["rmult"] = Y * (2 * U + 1)
    c["lmult"] = Y * (2 * U + 2)
    c["pinv"] = c["inv_gram"] + c["rmult"]
    c["project"] = c["lmult"] + c["pinv"] + (U + Y)
    c["inv_jacprod"] = Y * (2 * U * U + 3 * U + 8) + 2 * U * U * U // 3 + 2 * U * U \
        + Y * (2 * U + 8)
    Q = U + Y
    c["newton_iter"] = c["jacob"] + Y + c["inv_jacprod"] + c["rmult"] + 3 * Q
    c["quasi_newton_iter"] = c["constr"] + Y + c["pinv"] + 3 * Q
    c["grad_logdet"] = 2 * U * U * U + Y * (2 * U * U + U + 6) + c["second_order"]
    c["evaluate"] = c["jacob"] + c["decompose"] + (U + 2) * 30 + 3 * Q + c["grad_logdet"]
    c["step_fixed"] = 3 * c["project"] + 12 * Q + c["evaluate"]
    return c


def hh_counts(Y, n_steps_used, U=7, ND=6):
    """Hodgkin-Huxley (streaming family, forward-mode jets).  One exponential-Euler step
    of the recurrence as written in csrc/mlb_ode_hh.cuh has 21 jet*jet products (M), 21
    jet additions (A), 36 scalar-jet operations (S) and 24 unary functions exp / expm1 /
    reciprocal (C).  Per operation: value only M=A=S=C=1; first order (ND directions)
    M=1+3 ND, A=S=1+ND, C=2+ND; second order (NP=ND(ND+1)/2 pairs, counted ONCE: the
    kernel's per-group recomputation of the first-order part is not algorithmic work)
    M=1+3 ND+7 NP, A=S=1+ND+NP, C=3+ND+4 NP."""
    NP = ND * (ND + 1) // 2
    nM, nA, nS, nC = 21, 21, 36, 24
    step0 = nM + nA + nS + nC
    step1 = nM * (1 + 3 * ND) + (nA + nS) * (1 + ND) + nC * (2 + ND)
    step2 = (nM * (1 + 3 * ND + 7 * NP) + (nA + nS) * (1 + ND + NP)
             + nC * (3 + ND + 4 * NP))  # fmt: skip
    c = {}
    c["constr"] = n_steps_used * step0 + 3 * Y
    c["jacob"] = n_steps_used * step1 + 5 * Y
    c["second_order"] = n_steps_used * step2 + 4 * NP * Y
    c["decompose"] = Y * (U * (U + 1) + U + 8) + U * U * U // 3 + 20 * U
    c["inv_gram"] = Y * (2 * U + 8) + 2 * U * U + Y * (2 * U + 8)
    c["rmult"] = Y * (2 * U + 1)
    c["lmult"] = Y * (2 * U + 2)
    c["pinv"] = c["inv_gram"] + c["rmult"]
    Q = U + Y
    c["project"] = c["lmult"] + c["pinv"] + Q
    c["inv_jacprod"] = Y * (2 * U * U + 3 * U + 8) + 2 * U * U * U // 3 + 2 * U * U \
        + Y * (2 * U + 8)
    c["newton_iter"] = c["jacob"] + Y + c["inv_jacprod"] + c["rmult"] + 3 * Q
    c["quasi_newton_iter"] = c["constr"] + Y + c["pinv"] + 3 * Q
    c["grad_logdet"] = 2 * U * U * U + Y * (2 * U * U + U + 6) + c["second_order"]
    c["evaluate"] = c["jacob"] + c["decompose"] + (U + 2) * 30 + 3 * Q + c["grad_logdet"]
    c["step_fixed"] = 3 * c["project"] + 12 * Q + c["evaluate"]
    return c


def step_flops(model, solver, mean_iters_per_step):
    """Algorithmic flops of one constrained leapfrog step for one chain, given the mean
    number of projection iterations (forward + reverse) per step."""
    if model.model_id == 3:
        c = hh_counts(model.dim_y, model.n_steps_used)
    elif model.model_id == 2:
        c = fn_counts(model.dim_y, int(model.dt_seq.size))
    else:
        c = toy_counts() if model.model_id == 0 else hier_counts(model.dim_u, model.dim_y)
    per_iter = c["quasi_newton_iter"] if solver in (0, 4) else c["newton_iter"]
    return c["step_fixed"] + mean_iters_per_step * per_iter


def step_bytes(model):
    """Algorithmic HBM bytes of one chain-step: read + write pos, mom, plus
    16 B of per-chain outputs."""
    return 32 * model.dim_q + 16
