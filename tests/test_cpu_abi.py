"""CPU tests of the drop-in boundary (no GPU, no compute calls): libmlb.so loads, exports
every symbol include/mlb.h declares, the Python enums match the header, host-only entry
points behave, and the reference-facing Python surface has the reference's names,
keyword arguments and defaults."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "mlb.h")).read()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(mlb_[a-z0-9_]+)\s*\(", body)))


def test_library_exports_every_declared_symbol():
    from manifold_lifting_b200 import _lib

    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/mlb.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "manifold_lifting_b200._lib.EXPORTS out of sync"
    assert L.mlb_version() == int(re.search(r"#define MLB_VERSION (\d+)", HEADER).group(1))


def test_python_enums_match_header():
    from manifold_lifting_b200 import _lib

    def enum(name):
        m = re.search(rf"\b{name}\s*=\s*(-?\d+)", HEADER)
        assert m, name
        return int(m.group(1))

    for py, c in [("MODEL_TOY", "MLB_MODEL_TOY"), ("MODEL_HIER", "MLB_MODEL_HIER"),
                  ("MODEL_FN", "MLB_MODEL_FN"), ("MODEL_HH", "MLB_MODEL_HH"),
                  ("PARAM_NORMAL", "MLB_PARAM_NORMAL"),
                  ("SOLVER_QUASI_NEWTON", "MLB_SOLVER_QUASI_NEWTON"),
                  ("SOLVER_NEWTON", "MLB_SOLVER_NEWTON"),
                  ("SOLVER_NEWTON_LS", "MLB_SOLVER_NEWTON_LS"),
                  ("SOLVER_MICI_NEWTON", "MLB_SOLVER_MICI_NEWTON"),
                  ("SOLVER_MICI_QUASI_NEWTON", "MLB_SOLVER_MICI_QUASI_NEWTON"),
                  ("NORM_EUCLIDEAN", "MLB_NORM_EUCLIDEAN"), ("ST_DIVERGED", "MLB_ST_DIVERGED"),
                  ("ST_NOT_CONVERGED", "MLB_ST_NOT_CONVERGED"),
                  ("ST_NON_REVERSIBLE", "MLB_ST_NON_REVERSIBLE"),
                  ("ST_H_DIVERGED", "MLB_ST_H_DIVERGED"), ("ST_LINALG", "MLB_ST_LINALG")]:  # fmt: skip
        assert getattr(_lib, py) == enum(c), (py, c)
    assert _lib.ADAPT_ROWS == int(re.search(r"#define MLB_ADAPT_ROWS (\d+)", HEADER).group(1))
    assert _lib.STAT_ROWS == int(re.search(r"#define MLB_STAT_ROWS (\d+)", HEADER).group(1))
    for py, c in [("FLAG_SINGLE_KERNEL", "MLB_FLAG_SINGLE_KERNEL"),
                  ("FLAG_NUTS_NO_EXTRA_SUBTREE_CHECKS", "MLB_FLAG_NUTS_NO_EXTRA_SUBTREE_CHECKS"),
                  ("FLAG_NUTS_TWO_SIDED_DELTA_H", "MLB_FLAG_NUTS_TWO_SIDED_DELTA_H")]:
        assert getattr(_lib, py) == int(re.search(rf"#define {c} (\d+)", HEADER).group(1))
    assert C.sizeof(_lib.SolverOpts) == 56 and C.sizeof(_lib.ModelDesc) == 24


def test_model_handles_and_argument_errors():
    """Host-only entry points: model registry, dimensions, error codes (no CUDA calls)."""
    from manifold_lifting_b200 import _lib, models

    L = _lib.lib()
    m = models.EightSchoolsModel(np.arange(8.0), np.full(8, 10.0))
    assert (m.dim_u, m.dim_y, m.dim_q) == (2, 8, 18)
    assert L.mlb_model_jac_rows(m.handle) == 8 * 2 + 2 * 8
    assert L.mlb_model_gram_rows(m.handle) == 4 + 8 and L.mlb_model_gram_dim(m.handle) == 2
    t = models.Toy2DModel(0.1)
    assert (t.dim_u, t.dim_y, t.dim_q) == (2, 1, 3) and L.mlb_model_gram_dim(t.handle) == 1
    # unsupported shape: loud error, not a fallback
    with pytest.raises(_lib.MlbError):
        models.EightSchoolsModel(np.arange(7.0), np.full(7, 10.0))
    # bad arguments are reported as MLB_ERR_ARG with a message, before any CUDA work
    rc = L.mlb_constr(m.handle, C.c_int64(4), C.c_int64(2), None, None, None)
    assert rc == -1 and b"bad argument" in L.mlb_last_error()
    # n == 0 is a no-op success on every array entry
    assert L.mlb_constr(m.handle, C.c_int64(0), C.c_int64(0), None, None, None) == 0


def test_philox_host_entry_points_match_oracle():
    import c_oracle as co
    from manifold_lifting_b200 import _lib

    L = _lib.lib()
    for chain in (0, 5, 2**33 + 17):
        out = np.empty(18)
        L.mlb_philox_normals_host(1234, chain, 7, 18, out.ctypes.data)
        assert np.abs(out - co.philox_normals(1234, chain, 7, 18)).max() < 1e-14
        assert abs(L.mlb_philox_uniform_host(1234, chain, 7)
                   - co.philox_uniform(1234, chain, 7)) < 1e-16  # fmt: skip


def test_no_cpu_fallback_without_cuda():
    """The product path must fail loudly when there is no CUDA device."""
    import torch
    from manifold_lifting_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.MlbError):
        _lib.require_cuda()
    import manifold_lifting_b200 as mlb

    m = mlb.models.Toy2DModel(0.1)
    with pytest.raises(_lib.MlbError):  # constructing a system already demands the device
        mlb.IndependentAdditiveNoiseModelSystem(m, m.data, m.dim_u)
    with pytest.raises(_lib.MlbError):
        _lib.p_soa(torch.zeros(4, 3, dtype=torch.float64), 3)  # host tensors are refused


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: no file of the package may reference it."""
    pkg = os.path.join(ROOT, "manifold_lifting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f


def test_reference_api_surface():
    """Names / keyword arguments / defaults of mlift.solvers (solvers.py:16-26,98-108,
    172-183), mlift.systems classes (systems.py:508-517, 662-669) and `from mlift import *`
    (mlift/__init__.py:1-3)."""
    import manifold_lifting_b200 as mlb

    for name, extra in [("jitted_solve_projection_onto_manifold_quasi_newton", {}),
                        ("jitted_solve_projection_onto_manifold_newton", {}),
                        ("jitted_solve_projection_onto_manifold_newton_with_line_search",
                         {"max_line_search_iters": 10})]:  # fmt: skip
        fn = getattr(mlb, name)
        sig = inspect.signature(fn)
        assert list(sig.parameters)[:4] == ["state", "state_prev", "dt", "system"]
        want = dict(constraint_tol=1e-9, position_tol=1e-8, divergence_tol=1e10,
                    max_iters=50, norm=mlb.maximum_norm, **extra)  # fmt: skip
        for k, v in want.items():
            assert sig.parameters[k].default == v, (name, k)
    for cls in ("IndependentAdditiveNoiseModelSystem", "HierarchicalLatentVariableModelSystem"):
        sig = inspect.signature(getattr(mlb, cls).__init__)
        assert list(sig.parameters)[1:4] == ["generate_y", "data", "dim_u"]
    base = mlb.systems._AbstractDifferentiableGenerativeModelSystem
    for attr in ("_constr", "_jacob_constr_blocks", "_decompose_gram", "_log_det_sqrt_gram",
                 "_val_and_grad_log_det_sqrt_gram", "_lmult_by_jacob_constr",
                 "_rmult_by_jacob_constr", "_lmult_by_pinv_jacob_constr",
                 "_lmult_by_inv_jacob_product", "_normal_space_component",
                 "_quasi_newton_projection", "_newton_projection",
                 "_newton_projection_with_line_search", "precompile_jax_functions",
                 "h1", "dh1_dpos", "h2", "dh2_dmom", "h2_flow", "h1_flow", "h",
                 "project_onto_cotangent_space", "sample_momentum", "constr",
                 "jacob_constr_blocks", "gram_components", "log_det_sqrt_gram",
                 "grad_log_det_sqrt_gram", "normal_space_component"):  # fmt: skip
        assert hasattr(base, attr), attr
    assert hasattr(mlb, "ToleranceAdapter")
    sig = inspect.signature(mlb.ToleranceAdapter.__init__)
    assert sig.parameters["warm_up_constraint_tol"].default == 1e-6
    assert sig.parameters["warm_up_position_tol"].default == 1e-5
    me = mlb.install_as_mlift()
    import mlift
    import mlift.solvers
    import mlift.systems

    assert mlift is me and mlift.solvers is mlb.solvers
    # out-of-scope baselines stay importable by name (transitions.py:9,66,129)
    import mlift.transitions  # noqa: F401
    # mlift.linalg: all four functions with the reference's parameter names (linalg.py:6,12,46,88)
    import mlift.linalg as la

    for fn, params in [("lu_triangular_solve", ["l", "u", "b"]),
                       ("jvp_cholesky_mtx_mult_by_vct", ["tangents", "chol_mtx", "vct"]),
                       ("tridiagonal_solve", ["a", "b", "c", "d"]),
                       ("tridiagonal_pos_def_log_det", ["a", "b"])]:
        assert list(inspect.signature(getattr(la, fn)).parameters) == params, fn
    # closures of GaussianProcessModelSystem by the reference's names (systems.py:878-913)
    for fn, params in [("lmult_by_jacob_constr", ["dy_du", "dy_dn", "vct"]),
                       ("rmult_by_jacob_constr", ["dy_du", "dy_dn", "vct"]),
                       ("decompose_gram", ["dy_du", "dy_dn"]),
                       ("lmult_by_inv_gram", ["dy_du", "dy_dn", "chol_cap_mtx", "vct"]),
                       ("lmult_by_inv_jacob_product",
                        ["dy_du_l", "dy_dn_l", "dy_du_r", "dy_dn_r", "vct"])]:
        assert list(inspect.signature(getattr(mlb.gaussian_process, fn)).parameters) == params, fn
    # the dynamic sampler is built as the reference builds Mici's (utils.py:286-288) and
    # carries Mici's defaults
    sig = inspect.signature(mlb.samplers.DynamicMultinomialHMC.__init__)
    assert list(sig.parameters)[1:4] == ["system", "integrator", "rng"]
    assert sig.parameters["max_tree_depth"].default == 10
    assert sig.parameters["max_delta_h"].default == 1000.0
    assert sig.parameters["do_extra_subtree_checks"].default is True


def test_header_is_plain_c():
    """The boundary is a C ABI: include/mlb.h must compile as C99 (no C++ in signatures)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    header = os.path.join(ROOT, "include", "mlb.h")
    for std, lang in (("-std=c99", "c"), ("-std=c++17", "c++")):
        r = subprocess.run([gcc, std, "-fsyntax-only", "-x", lang, header],
                           capture_output=True, text=True)  # fmt: skip
        assert r.returncode == 0, r.stderr
