"""GPU suite for the LARGE-component path (shared-memory tile kernels, pasif_b200/csrc/large_kernels.cuh), called
through the C ABI:
  * the reference's golden vectors again, with every component forced through the tile kernels (large_threshold=0);
  * oscillator chains (BASELINE config 4 shape, scaled down) against the C oracle: sparse row programs and the
    tensor-core GEMM path (dense shared K'), forward + trajectory + adjoint, both schedules, both precisions;
  * several instances with their own coefficients, ragged tiles (E not a multiple of the tile width).
Tolerances as in test_parity_gpu.py (fp64 <= 1e-12 per variable; fp32 <= 1e-5 on the whole vector, forward and adjoint)."""
import numpy as np
import pytest

from conftest import golden, rel_err, col_err, run_case
from pasif_b200 import capi, workloads as wl
from pasif_oracle_c import CRecordingDriver

pytestmark = pytest.mark.gpu

TOL_FWD = {"f64": 1e-12, "f32": 1e-5}
TOL_ADJ = {"f64": 1e-12, "f32": 1e-5}
FORCE_LARGE = {"large_threshold": "0", "lanes": "off"}


def _oracle(prec, compat):
    class D(CRecordingDriver):
        dtype = np.float64 if prec == "f64" else np.float32
    D.compat = compat
    return D


def close(a, b, width, prec, tol):
    if prec == "f64":
        return col_err(a, b, width) < tol[prec]
    return rel_err(a, b) < tol[prec] and col_err(a, b, width) < 2e-3


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["mixed2", "fit8x2_jit", "ex1dof_short", "ex1dof_sq", "ex1dof_interp", "fit8_e16"])
def test_tile_kernels_match_reference_golden(name, prec):
    z = golden(name, prec)
    case = wl.golden_case(name)
    out = run_case(capi.engine_class(prec), case, options=FORCE_LARGE)
    assert out["_p"]._query("large_classes") >= 1
    E = len(case["excitations"])
    assert close(out["amplitudes"], z["amplitudes"], z["amplitudes"].size // E, prec, TOL_FWD)
    assert np.array_equal(out["amplitudes"], out["amplitudes_2nd"])
    if "trajectory" in z.files:
        assert out["trajectory"].shape == z["trajectory"].shape
        assert rel_err(out["trajectory"][1:], z["trajectory"][1:]) < TOL_FWD[prec]
    if "gradient" in z.files:
        assert close(out["gradient"], z["gradient"], z["gradient"].size // E, prec, TOL_ADJ)


def chain_case(ndof, dense, E, L, n_sys=1, seed=3):
    rng = np.random.default_rng(seed)
    base = wl.chain_system(ndof, dense=dense, seed=seed)
    systems = [base] if n_sys == 1 else [wl.jitter(base, rng, rel=0.05) for _ in range(n_sys)]
    return dict(systems=systems, jacobians=[wl.chain_jacobian(s, 3) for s in systems],
                excitations=wl.random_sine_excitations(E, L, 16000, seed=seed + 1), sample_rate=16000, numsteps=0,
                interp=None, traj_save_steps=max(L // 7, 1))


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("compat", ["reference", "clean"])
@pytest.mark.parametrize("mode", ["sparse", "dense-stage", "dense-tile"])
def test_chain_matches_oracle(mode, compat, prec):
    """sparse: row programs in the tile kernels (lane-block split switched off); dense-stage: per-stage batched GEMM +
    fused update with the state in HBM (the default for a dense shared K'); dense-tile: GEMM inside the tile kernels."""
    dense = mode != "sparse"
    case = chain_case(40, dense, E=11, L=700)
    a = run_case(capi.engine_class(prec, compat), case, options={"gemm": "tile"} if mode == "dense-tile" else ({"lanes": "off"} if mode == "sparse" else None))
    b = run_case(_oracle(prec, compat), case)
    assert a["_p"]._query("large_classes") >= 1
    assert a["_p"]._query("lane_classes") == 0
    assert (a["_p"]._query("large_stage_classes") >= 1) == (mode == "dense-stage")
    assert (a["_p"]._query("large_gemm_classes") >= 1) == dense
    n, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    assert close(a["amplitudes"], b["amplitudes"], n, prec, TOL_FWD)
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)


def test_gemm_path_is_used_and_agrees_with_row_programs():
    """Dense shared K': the tensor-core path (DMMA) and the sparse row programs (gemm=off) give the same fp64 result."""
    case = chain_case(48, True, E=16, L=400)
    b = run_case(capi.engine_class("f64"), case, options={"gemm": "off", "lanes": "off"})
    assert b["_p"]._query("large_gemm_classes") == 0
    n, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    for mode in ("stage", "tile"):
        a = run_case(capi.engine_class("f64"), case, options={"gemm": mode})
        assert a["_p"]._query("large_gemm_classes") >= 1
        assert col_err(a["amplitudes"], b["amplitudes"], n) < 1e-12
        assert rel_err(a["trajectory"], b["trajectory"]) < 1e-12
        assert col_err(a["gradient"], b["gradient"], na) < 1e-12


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_several_chain_instances_ragged_tiles(prec):
    case = chain_case(30, False, E=19, L=500, n_sys=3)
    a = run_case(capi.engine_class(prec), case, options={"lanes": "off", "call_order": "independent"})
    b = run_case(_oracle(prec, "reference"), case)
    n = sum(s["n"] for s in case["systems"])
    na = sum(s["n"] for s in case["jacobians"])
    assert close(a["amplitudes"], b["amplitudes"], n, prec, TOL_FWD)
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)
    red = np.asarray(a["_p"]._getGradientReduced(0))
    assert rel_err(red, a["gradient"].reshape(19, na).sum(axis=0)) < (1e-12 if prec == "f64" else 1e-5)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_stage_path_two_overlapping_parts(prec):
    """> 256 trajectories: the dense-operator path splits the batch into two parts on two streams (GEMM of one part
    overlaps the update of the other); the result is the same as with overlap=off (bit for bit) and matches the oracle."""
    case = chain_case(32, True, E=300, L=60)
    a = run_case(capi.engine_class(prec), case)
    s = run_case(capi.engine_class(prec), case, options={"overlap": "off"})
    b = run_case(_oracle(prec, "reference"), case)
    assert a["_p"]._query("large_stage_classes") >= 1
    n, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    for key in ("amplitudes", "trajectory", "gradient"):
        assert np.array_equal(a[key], s[key]), key
    assert close(a["amplitudes"], b["amplitudes"], n, prec, TOL_FWD)
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)


@pytest.mark.parametrize("shape", [(100, 300), (514, 256), (66, 1000)])
def test_stage_gemm_kernels_against_numpy(shape):
    """The three GEMM kernels of the dense-operator path in isolation: DMMA is exact to fp64 round-off; both 3xTF32
    kernels (mma.sync and tcgen05/TMEM/TMA) reach fp32-level accuracy and agree with each other."""
    N, P = shape
    rng = np.random.default_rng(N)
    kd, x = rng.standard_normal((N, N)), rng.standard_normal((P, N))
    ref = x @ kd.T
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(capi.debug_gemm(kd, x, "f64", 0) - ref)) < 1e-13 * scale
    y_mma, y_tc, y_fs = capi.debug_gemm(kd, x, "f32", 0), capi.debug_gemm(kd, x, "f32", 1), capi.debug_gemm(kd, x, "f32", 2)
    assert np.max(np.abs(y_mma - ref)) < 1e-5 * scale
    assert np.max(np.abs(y_tc - ref)) < 1e-5 * scale
    assert np.max(np.abs(y_tc - y_mma)) < 2e-6 * scale
    assert np.array_equal(y_fs, y_tc)          # fused in-kernel split: same tf32 halves, same products, same order


@pytest.mark.parametrize("shape", [(256, 300), (512, 1024)])
def test_wide_tile_gemm_is_bit_identical(shape):
    """256-row tiles of the fused-split tcgen05 GEMM (the default when the band of dense rows is a multiple of 256, as in
    BASELINE config 4): same products in the same order as the 128-row tiles."""
    N, P = shape
    rng = np.random.default_rng(N + P)
    kd, x = rng.standard_normal((N, N)), rng.standard_normal((P, N))
    ref = x @ kd.T
    y_fs, y_wide = capi.debug_gemm(kd, x, "f32", 2), capi.debug_gemm(kd, x, "f32", 4)
    assert np.max(np.abs(y_wide - ref)) < 1e-5 * np.max(np.abs(ref))
    assert np.array_equal(y_fs, y_wide)
    # 16-float k-blocks (64-byte swizzle, twice the pipeline stages): the same products in the same order
    assert np.array_equal(capi.debug_gemm(kd, x, "f32", 5), y_fs)
    assert np.array_equal(capi.debug_gemm(kd, x, "f32", 6), y_fs)
    # CTA pairs (tcgen05.mma.cta_group::2, M = 256 trajectories, every CTA holds half of the K' tile): one 256-row tile per
    # pair, and two tiles that share the X tile when the rows are a multiple of 512
    assert np.array_equal(capi.debug_gemm(kd, x, "f32", 7), y_fs)
    if N % 512 == 0:
        assert np.array_equal(capi.debug_gemm(kd, x, "f32", 8), y_fs)


def test_config4_size_dense_chain_matches_oracle_f32():
    """The 256-DOF dense chain of BASELINE config 4 (514 states, band of 256 dense rows -> 256-row GEMM tiles) on a few
    excitations: forward and gradient against the oracle, and bit-identical to the 128-row tiles (tc_bn=128)."""
    case = chain_case(256, True, E=12, L=24)
    a = run_case(capi.engine_class("f32"), case)
    n = run_case(capi.engine_class("f32"), case, options={"tc_bn": "128"})
    b = run_case(_oracle("f32", "reference"), case)
    assert a["_p"]._query("large_stage_classes") >= 1
    assert a["_p"]._query("stage_band_rows_fwd") == 256
    nf, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    for key in ("amplitudes", "trajectory", "gradient"):
        assert np.array_equal(a[key], n[key]), key
    assert close(a["amplitudes"], b["amplitudes"], nf, "f32", TOL_FWD)
    assert close(a["gradient"], b["gradient"], na, "f32", TOL_ADJ)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_config4_full_size_dense_chain_128_excitations(prec):
    """BASELINE config 4 at its full state size (256-DOF dense chain: 514 states, 518-state adjoint) on 128 excitations --
    one whole 128-trajectory GEMM tile -- in both precisions: amplitudes and gradient against the oracle."""
    case = chain_case(256, True, E=128, L=24)
    case["traj_save_steps"] = None
    a = run_case(capi.engine_class(prec), case)
    b = run_case(_oracle(prec, "reference"), case)
    assert a["_p"]._query("large_stage_classes") >= 1 and a["_p"]._query("stage_band_rows_fwd") == 256
    nf, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    assert close(a["amplitudes"], b["amplitudes"], nf, prec, TOL_FWD)
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)
    assert np.max(np.abs(a["gradient"])) > 0


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("E", [48, 50])
def test_parameter_replicas_on_the_dense_operator_path(E, prec):
    """BASELINE config 4's "parameter replicas": several instances that share the dense linear operator K' bit for bit and
    differ in their Gamma / Lambda / F values (workloads.jitter_nonlinear).  They form ONE class on the dense-operator
    path (one GEMM for all of them, per-instance coefficients in the update kernel); results against the oracle, which
    integrates the block-diagonal global system."""
    base = wl.chain_system(40, dense=True, seed=3)
    rng = np.random.default_rng(11)
    systems = [wl.jitter_nonlinear(base, rng) for _ in range(3)]
    case = dict(systems=systems, jacobians=[wl.chain_jacobian(s, 3) for s in systems],
                excitations=wl.random_sine_excitations(E, 300, 16000, seed=6), sample_rate=16000, numsteps=0, interp=None,
                traj_save_steps=60)
    a = run_case(capi.engine_class(prec), case)
    b = run_case(_oracle(prec, "reference"), case)
    p = a["_p"]
    assert p._query("large_stage_classes") >= 1 and p._query("large_gemm_classes") >= 1      # (forward class + joint class)
    assert p._query("joint_classes") == 1 and p._query("joint_instances") == 3
    nf, na = sum(s["n"] for s in systems), sum(j["n"] for j in case["jacobians"])
    assert close(a["amplitudes"], b["amplitudes"], nf, prec, TOL_FWD)
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)
    g = a["gradient"].reshape(E, 3, -1)
    assert not np.array_equal(g[:, 0], g[:, 1])                # the replicas really differ


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("variant", ["gamma_offsite", "lambda_ab_offsite", "lambda_c_offsite"])
def test_dense_operator_offsite_nonlinear_terms(variant, prec):
    """The update kernel of the dense-operator path shares one operand gather between the operands of a term when their
    indices coincide in EVERY entry of the class (on-site x^2 / x^3).  Here some do not: x_i x_{i+1}, x_i x_{i+1}^2,
    x_i^2 x_{i+1} couplings clear one flag each, so every combination of the shared / separate gathers is compared with
    the oracle, forward and adjoint (Jacobian from pasif_b200.adjoint.build_adjoint)."""
    base = wl.chain_system(40, dense=True, seed=3)
    n = 40
    s = dict(base)
    gd, gv, gi = base["Gamma"]
    ld, lv, li = base["Lambda"]
    gi, li = list(gi), list(li)
    for i in range(n - 1):
        if variant == "gamma_offsite":
            gi[3 * (i + 1) + 2] = i + 1                     # [n + i, i, i + 1]   (entry 0 is the probe term)
        elif variant == "lambda_ab_offsite":
            li[4 * i + 2], li[4 * i + 3] = i + 1, i + 1     # [n + i, i, i + 1, i + 1]
        else:
            li[4 * i + 3] = i + 1                           # [n + i, i, i, i + 1]
    s["Gamma"], s["Lambda"] = (gd, gv, gi), (ld, lv, li)
    case = dict(systems=[s], jacobians=[wl.chain_jacobian(s, 3)], excitations=wl.random_sine_excitations(12, 400, 16000, seed=8),
                sample_rate=16000, numsteps=0, interp=None, traj_save_steps=80)
    a = run_case(capi.engine_class(prec), case)
    b = run_case(_oracle(prec, "reference"), case)
    assert a["_p"]._query("large_stage_classes") >= 1
    nf, na = s["n"], case["jacobians"][0]["n"]
    assert close(a["amplitudes"], b["amplitudes"], nf, prec, TOL_FWD)
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("shape", ["small_ragged", "two_parts", "replicas_nonuniform", "config4_rows", "many_tiles"])
def test_stream_update_is_bit_identical(shape, prec):
    """The persistent bulk-copy-pipelined update kernel (stage_stream.cuh, the default) does the arithmetic of the
    per-thread update kernel (update_stream=off) term by term in the same order: every output is equal bit for bit --
    ragged batches (E not a multiple of the tile of trajectories), several parts on their own streams, instances with
    their own coefficients whose boundaries fall inside a tile, and the full 514-state system of BASELINE config 4;
    tiles of 2, 4 and 8 trajectories, one or two CTAs per SM."""
    if shape == "small_ragged":
        case = chain_case(40, True, E=11, L=200)
    elif shape == "two_parts":
        case = chain_case(32, True, E=301, L=40)
    elif shape == "replicas_nonuniform":
        base = wl.chain_system(40, dense=True, seed=3)
        rng = np.random.default_rng(11)
        systems = [wl.jitter_nonlinear(base, rng) for _ in range(3)]
        case = dict(systems=systems, jacobians=[wl.chain_jacobian(s, 3) for s in systems],
                    excitations=wl.random_sine_excitations(50, 120, 16000, seed=6), sample_rate=16000, numsteps=0, interp=None,
                    traj_save_steps=40)
    elif shape == "many_tiles":
        # several tiles per CTA: the buffer ring of the kernel wraps (the grid holds at most 2 CTAs per SM)
        case = chain_case(32, True, E=4800, L=6)
        case["traj_save_steps"] = None
    else:
        case = chain_case(256, True, E=130, L=16)
        case["traj_save_steps"] = None
    ref = run_case(capi.engine_class(prec), case, options={"update_stream": "off"})
    assert ref["_p"]._query("large_stage_classes") >= 1
    for tt, ctas in (("0", "2"), ("2", "2"), ("4", "2"), ("0", "1"), ("8", "1")):
        if shape == "config4_rows" and tt == "2":
            continue
        a = run_case(capi.engine_class(prec), case, options={"update_stream": "on", "stream_tt": tt, "stream_ctas": ctas})
        for key in ("amplitudes", "trajectory", "gradient"):
            if key in ref and ref[key] is not None:
                assert np.array_equal(a[key], ref[key]), (key, tt)


def test_gemm_launch_variants_are_bit_identical():
    """The fp32 stage GEMM has several launch forms that run the same products in the same order: CTA pairs
    (tcgen05.mma.cta_group::2, the default when the part is a multiple of 256 trajectories and the band a multiple of 256
    rows; two row tiles per pair for the 512-row adjoint band) or single CTAs, programmatic dependent launch on / off, fewer
    pipeline stages with a shared-memory budget for the update kernel that leaves room for a GEMM CTA on the same SM, parts
    that start one GEMM apart, tiles of the update kernel drawn from a counter instead of a fixed share per CTA.  Amplitudes and gradient of the full 514-state chain are equal bit for bit in all of them."""
    case = chain_case(256, True, E=512, L=6)
    case["traj_save_steps"] = None
    ref = run_case(capi.engine_class("f32"), case, options={"tc_pair": "off", "pdl": "off"})
    assert ref["_p"]._query("large_stage_classes") >= 1
    try:
        for opts in ({"tc_pair": "on", "pdl": "on"},
                     {"tc_pair": "on", "tc_pair_nt": "1", "pdl": "off"},
                     {"tc_pair": "off", "pdl": "on", "stage_parts": "1"},
                     {"tc_pair": "off", "tc_stages": "3", "stream_smem_kb": "80", "stream_grid": "1", "stage_dephase": "on", "stage_parts": "4"},
                     {"tc_pair": "off", "tc_stages": "2", "stream_smem_kb": "128", "stream_grid": "1", "stage_parts": "2"},
                     {"stream_dynamic": "on", "stream_grid_ctas": "168", "stage_dephase": "on"},
                     {"stream_dynamic": "on", "stage_parts": "3"}):
            a = run_case(capi.engine_class("f32"), case, options=opts)
            for key in ("amplitudes", "gradient"):
                assert np.array_equal(a[key], ref[key]), (key, opts)
    finally:
        run_case(capi.engine_class("f32"), chain_case(8, True, E=4, L=4), options={"pdl": "on"})   # (process-wide switch back to its default)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_dense_operator_exact_schedule_in_one_chunk(prec):
    """compat=clean on the dense-operator path: when the forward states of all steps fit the free device memory the run is
    ONE chunk (one forward sweep that writes every state, no checkpoints, no recompute); with tail_budget_gb=0 it is the
    square-root schedule (sweep with checkpoints + recompute of every chunk).  Recomputed states equal swept states, so the
    gradients are equal bit for bit -- and both match the oracle's exact schedule."""
    case = chain_case(40, True, E=24, L=90)
    one = run_case(capi.engine_class(prec, "clean"), case)
    sq = run_case(capi.engine_class(prec, "clean"), case, options={"tail_budget_gb": "0"})
    assert one["_p"]._query("large_stage_classes") >= 1
    assert one["_p"]._query("stored_chunks") == one["_p"]._query("chunks") == 9        # 89 steps in chunks of ceil(sqrt 89) = 10
    assert sq["_p"]._query("stored_chunks") == 0
    assert np.array_equal(one["gradient"], sq["gradient"])
    b = run_case(_oracle(prec, "clean"), case)
    assert close(one["gradient"], b["gradient"], case["jacobians"][0]["n"], prec, TOL_ADJ)


def test_fp64_gemm_tile_heights_are_bit_identical():
    """fp64 stage GEMM (DMMA): 32-row tiles (default: four CTAs of four warps per SM, so the SMs finish within one small tile of
    each other) and 64-row tiles accumulate every element over k in the same order -- gradients equal bit for bit."""
    case = chain_case(256, True, E=130, L=6)
    case["traj_save_steps"] = None
    a = run_case(capi.engine_class("f64"), case, options={"dg_bm": "64"})
    b = run_case(capi.engine_class("f64"), case)
    assert a["_p"]._query("large_stage_classes") >= 1
    for key in ("amplitudes", "gradient"):
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("compat", ["reference", "clean"])
def test_dense_operator_gradient_graph_replay(compat, prec):
    """The launch sequence of a dense-operator gradient is captured into a CUDA graph on the second call with the same
    configuration and replayed from the third on: every call returns the same bits as the directly queued launches
    (stage_graph=off), also after new excitations of the same shape have been loaded."""
    case = chain_case(40, True, E=24, L=60)
    a = run_case(capi.engine_class(prec, compat), case)
    p = a["_p"]
    assert p._query("large_stage_classes") >= 1
    g = [a["gradient"]] + [np.asarray(p.getGradient()) for _ in range(3)]
    assert p._query("stage_graph_replays") >= 2
    for x in g[1:]:
        assert np.array_equal(x, g[0])
    off = run_case(capi.engine_class(prec, compat), case, options={"stage_graph": "off"})
    assert off["_p"]._query("stage_graph_replays") == 0
    assert np.array_equal(off["gradient"], g[0])
    # new excitations of the same shape: the graph reads the refreshed device buffer
    case2 = dict(case, excitations=wl.random_sine_excitations(24, 60, 16000, seed=99))
    p.setExcitations([np.array(r) for r in case2["excitations"]], case2["sample_rate"])
    g2 = np.asarray(p.getGradient())
    ref2 = run_case(capi.engine_class(prec, compat), case2, options={"stage_graph": "off"})
    assert np.array_equal(g2, ref2["gradient"])
    assert not np.array_equal(g2, g[0])
