"""CPU-only tests: host logic, C-ABI surface, statistics, sharding/gather over gloo, Philox oracle."""
import argparse
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import nested_diffusion_b200 as nd
from nested_diffusion_b200 import _capi, schedule, stats
from oracle import ladine_oracle as orc
from oracle import philox_oracle as pho
from tests.golden_util import Fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI surface
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ladine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ladine_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libladine.so lacks {name} declared in include/ladine.h"
    assert sorted(_capi.SYMBOLS) == declared, "ctypes binding and header disagree"
    assert lib.ladine_version() == 3


def test_header_is_plain_c_and_a_c_caller_links(tmp_path):
    """include/ladine.h must be consumable by a C compiler (the boundary is a C ABI: cgo / JNI / ctypes bind it), and a C
    translation unit that references every declared function must link against the shared library (no compute calls)."""
    import shutil

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(ROOT, "include", "ladine.h")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    names = sorted(set(re.findall(r"\b(ladine_[a-z0-9_]+)\s*\(", open(hdr).read())))
    assert set(names) == set(_capi.SYMBOLS), set(names) ^ set(_capi.SYMBOLS)
    src = tmp_path / "caller.c"
    src.write_text('#include "ladine.h"\n#include <stdio.h>\nint main(void) {\n  const void* fns[] = {' +
                   ", ".join(f"(const void*){n}" for n in names) +
                   '};\n  printf("%d %d\\n", ladine_version(), (int)(sizeof fns / sizeof fns[0]));\n  return 0;\n}\n')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_capi.lib_path())
    r = subprocess.run([gcc, "-std=gnu99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                        "-l:libladine.so", f"-Wl,-rpath,{libdir}", "-Wl,--allow-shlib-undefined"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    if r.returncode == 0:      # libcudart may be unresolvable outside the torch process on a CPU-only box: linking is the test
        assert r.stdout.split() == [str(_capi.load().ladine_version()), str(len(names))]


def test_struct_sizes_match_header_layout():
    # sizes the C side checks through struct_size: any drift is an error at call time, not corruption
    assert ctypes.sizeof(_capi.MemberDesc) == 8 * 4 + 8 * 8 + 15 * 8
    assert ctypes.sizeof(_capi.SampleArgs) % 8 == 0


def test_build_freshness_is_content_based(monkeypatch):
    """The in-tree library is rebuilt when the sources or flags change, not when file times do (a snapshot copy to the
    GPU box does not preserve mtimes)."""
    from nested_diffusion_b200 import build as b

    b.build()
    assert b.up_to_date()
    h0 = b.source_hash()
    os.utime(b.SRC[0], None)                      # touching a source changes nothing
    assert b.up_to_date() and b.source_hash() == h0
    monkeypatch.setattr(b, "FLAGS", b.FLAGS + ["-DX"])
    assert not b.up_to_date()


def _plan(K, rows, Fp, geometry, row_major, units):
    lib = _capi.load()
    info = (ctypes.c_int32 * 4)()
    n = lib.ladine_debug_plan(K, rows, Fp, geometry, row_major, units, None, 0, info)
    assert n > 0
    buf = (ctypes.c_int32 * n)()
    assert lib.ladine_debug_plan(K, rows, Fp, geometry, row_major, units, buf, n, info) == n
    used, stride, rows_pad, row_tiles = list(info)
    table = np.frombuffer(buf, dtype=np.int32).reshape(used, stride)
    return table, rows_pad, row_tiles


@pytest.mark.parametrize("K,rows,Fp,geometry,units", [
    (5, 1400, 4096, 1, 148),     # config 2, single-CTA tiles
    (5, 1400, 4096, 2, 74),      # config 2 as CTA pairs: 5 full pair tiles + one half tile per (member, N tile)
    (1, 64, 4096, 3, 148),       # config 1, slim tiles
    (5, 20480, 4096, 2, 74),     # config 3
    (3, 300, 512, 1, 148), (2, 257, 1024, 2, 74), (8, 1, 256, 1, 148), (1, 129, 768, 3, 7),
])
@pytest.mark.parametrize("row_major", [0, 1])
def test_static_tile_schedule_covers_every_tile_once_and_balances(K, rows, Fp, geometry, units, row_major):
    """ladine_debug_plan (host-only): the schedule every GEMM launch reads.  Every (member, N tile, row tile) appears
    exactly once, rows terminate with -1, half tiles only where a pair geometry has <= 128 trailing rows, and the
    per-unit load is balanced to within one tile."""
    table, rows_pad, row_tiles = _plan(K, rows, Fp, geometry, row_major, units)
    tile_cols = 128 if geometry == 3 else 256
    tile_rows = 256 if geometry == 2 else 128
    NB = Fp // tile_cols
    rem = rows % tile_rows
    has_half = geometry == 2 and 0 < rem <= 128
    assert row_tiles == (rows // tile_rows) + (1 if rem else 0)
    assert rows_pad == row_tiles * tile_rows and rows_pad >= rows
    seen, loads, order = set(), [], []
    for unit in table:
        live = unit[unit >= 0]
        assert (unit[len(live):] == -1).all() and len(live) < len(unit), "tiles first, then the -1 terminator"
        cost = 0
        for code in live:
            member, nb, mb, half = int(code) >> 23, (int(code) >> 13) & 1023, (int(code) >> 1) & 4095, int(code) & 1
            assert 0 <= member < K and 0 <= nb < NB and 0 <= mb < row_tiles
            assert half == (1 if has_half and mb == row_tiles - 1 else 0)
            assert (member, nb, mb) not in seen
            seen.add((member, nb, mb))
            order.append((member, nb, mb))
            cost += 17 if half else 20
        loads.append(cost)
    assert len(seen) == K * NB * row_tiles
    assert len(table) == min(units, K * NB * row_tiles)
    assert max(loads) - min(loads) <= 20, "longest-processing-time quotas: no unit is more than one tile ahead"
    # neighbours in the canonical order run at the same time: the first tile of every unit belongs to one member and,
    # in N-tile-major order, to as few N tiles as possible (they share the W tile)
    first = [(int(u[0]) >> 23, (int(u[0]) >> 13) & 1023, (int(u[0]) >> 1) & 4095) for u in table]
    if len(table) <= NB * row_tiles:
        assert len({f[0] for f in first}) == 1
        if row_major:
            assert len({f[2] for f in first}) <= -(-len(table) // NB) + 1
        else:
            assert len({f[1] for f in first}) <= -(-len(table) // row_tiles) + 1


@pytest.mark.parametrize("K,rows,want,why", [
    (5, 20480, 2, "config 3: 87 single rounds vs 87 pair rounds, each 1/1.08 as long"),
    (5, 2560, 2, "an 8-GPU shard of config 3"),
    (5, 1400, 1, "config 2: 5.95 rounds of single tiles vs 6.7 pair rounds (measured 205 vs 210 us)"),
    (1, 1400, 2, "one member x 1400 rows (draws-ahead): 176 single tiles = 2 rounds on 148 SMs, 96 pair units = 2 shorter ones"),
    (1, 64, 3, "config 1: 16 wide tiles -> slim tiles still fit one round"),
    (1, 128, 3, "one row tile"),
    (2, 128, 3, "two members x one row tile: 32 wide tiles, 64 slim ones"),
    (5, 128, 1, "80 wide tiles: two rounds of slim ones -> wide"),
])
def test_auto_geometry_choice(K, rows, want, why):
    lib = _capi.load()
    assert lib.ladine_debug_geometry(K, rows, 4096, 148) == want, why
    assert lib.ladine_debug_geometry(0, rows, 4096, 148) < 0 and lib.ladine_debug_geometry(K, rows, 100, 148) < 0


def test_debug_plan_rejects_bad_arguments():
    lib = _capi.load()
    info = (ctypes.c_int32 * 4)()
    for args in ((0, 10, 256, 1, 0, 8), (1, 0, 256, 1, 0, 8), (1, 10, 300, 1, 0, 8), (1, 10, 256, 4, 0, 8),
                 (1, 10, 256, 1, 0, 0), (9, 10, 256, 1, 0, 8)):
        assert lib.ladine_debug_plan(*args, None, 0, info) < 0


def test_create_fails_cleanly_without_a_device():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = ctypes.c_void_p()
    rc = _capi.load().ladine_create(0, ctypes.byref(out))
    assert rc < 0 and not out.value


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nested_diffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_no_cpu_fallback():
    fx = Fixture("state_dict_layout")
    cfg = argparse.Namespace(diffusion=argparse.Namespace(timesteps=5),
                             data=argparse.Namespace(num_classes=2, dataset="ChestXRay"),
                             model=argparse.Namespace(data_dim=10, arch="linear", feature_dim=8, hidden_dim=6))
    m = nd.ConditionalModel(cfg, guidance=True).eval()
    a, o = schedule.schedule_tensors(schedule.make_beta_schedule("linear", 5, 1e-4, 0.02))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nd.diffusion_utils.p_sample_loop(m, fx["guid_c2/x"], fx["guid_c2/yh"], fx["guid_c2/yh"], 5, a, o)


# ------------------------------------------------------------------ drop-in module layout
@pytest.mark.parametrize("tag,guidance,C", [("guid_c2", True, 2), ("noguid_c3", False, 3)])
def test_conditional_model_matches_reference_layout_and_forward(tag, guidance, C):
    fx = Fixture("state_dict_layout")
    m = fx.meta
    cfg = argparse.Namespace(diffusion=argparse.Namespace(timesteps=m["T"]),
                             data=argparse.Namespace(num_classes=C, dataset="ChestXRay"),
                             model=argparse.Namespace(data_dim=m["Dx"], arch="linear", feature_dim=m["F"],
                                                      hidden_dim=m["H"]))
    model = nd.ConditionalModel(cfg, guidance=guidance).eval()
    mine = {k: [list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()}
    assert mine == m["layouts"][tag], "state_dict layout differs from the reference (checkpoint contract)"
    sd = {k[len(tag) + 4:]: v for k, v in fx.arrays.items() if k.startswith(tag + "/sd/")}
    model.load_state_dict(sd)  # strict
    with torch.no_grad():
        out = model(fx[f"{tag}/x"], fx[f"{tag}/y"], torch.tensor([3]), fx[f"{tag}/yh"])
        out_o = orc.denoiser_forward(sd, fx[f"{tag}/x"], fx[f"{tag}/y"], torch.tensor([3]), fx[f"{tag}/yh"])
    assert torch.allclose(out, fx[f"{tag}/out"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(out_o, fx[f"{tag}/out"], rtol=1e-6, atol=1e-7)


def test_drop_in_module_surface():
    du = nd.diffusion_utils
    for name in ("make_beta_schedule", "extract", "q_sample", "p_sample", "p_sample_t_1to0", "y_0_reparam",
                 "p_sample_loop"):
        assert callable(getattr(du, name))
    import inspect

    sig = inspect.signature(du.p_sample_loop)
    assert list(sig.parameters)[:10] == ["model", "x", "y_0_hat", "y_T_mean", "n_steps", "alphas",
                                         "one_minus_alphas_bar_sqrt", "only_last_sample",
                                         "input_model_original_version", "output_detach"]
    assert sig.parameters["only_last_sample"].default is False
    assert list(inspect.signature(du.p_sample).parameters)[:9] == [
        "model", "x", "y", "y_0_hat", "y_T_mean", "t", "alphas", "one_minus_alphas_bar_sqrt", "output_detach"]


# ------------------------------------------------------------------ schedules / coefficients
def test_module_swap_shims_bind_the_drop_in():
    """shims/ ahead of the reference on sys.path: the runner's own import lines resolve to this package."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from diffusion_utils import *\n"
            "from latent_model import ConditionalModel\n"
            "import nested_diffusion_b200.diffusion_utils as du, nested_diffusion_b200.latent_model as lm\n"
            "assert p_sample_loop is du.p_sample_loop and p_sample is du.p_sample and make_beta_schedule is du.make_beta_schedule\n"
            "assert ConditionalModel is lm.ConditionalModel\n" % (ROOT, os.path.join(ROOT, "shims")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_schedules_and_coef_table_equal_oracle_bitwise():
    fx = Fixture("schedules")
    for key, ref in fx.arrays.items():
        kind, T = key.split("/")
        got = schedule.make_beta_schedule(kind, int(T), fx.meta["start"], fx.meta["end"]).float()
        assert torch.allclose(got, ref, rtol=1e-6, atol=0), key
    for kind in ("linear", "cosine", "quad"):
        betas = schedule.make_beta_schedule(kind, 1000, 1e-4, 0.02)
        a, o = schedule.schedule_tensors(betas, kind)
        a2, o2 = orc.schedule_tensors(betas, kind)
        assert torch.equal(a, a2) and torch.equal(o, o2)
        assert torch.equal(schedule.coef_table(a, o, 1000), orc.coef_table(a, o, 1000))
        assert torch.equal(schedule.coef_table(a, o, 100), orc.coef_table(a, o, 100))


def test_coef_table_reproduces_reference_step_scalars():
    """row t of the table == the scalars diffusion_utils.p_sample computes at that t (same torch expressions)."""
    a, o = schedule.schedule_tensors(schedule.make_beta_schedule("linear", 50, 1e-4, 0.02))
    tab = schedule.coef_table(a, o, 50)
    y = torch.zeros(1, 2)
    for t in (49, 7, 1):
        tt = torch.tensor([t])
        alpha_t, s_t, s_p = orc.extract(a, tt, y), orc.extract(o, tt, y), orc.extract(o, tt - 1, y)
        q, qp = (1 - s_t.square()).sqrt(), (1 - s_p.square()).sqrt()
        g0 = (1 - alpha_t) * qp / s_t.square()
        g1 = s_p.square() * alpha_t.sqrt() / s_t.square()
        g2 = 1 + (q - 1) * (alpha_t.sqrt() + qp) / s_t.square()
        sig = (s_p.square() / s_t.square() * (1 - alpha_t)).sqrt()
        want = torch.stack([(1 / q), 1 - q, s_t, g0, g1, g2, sig]).flatten()
        assert torch.equal(tab[t, :7], want)


# ------------------------------------------------------------------ statistics
def test_statistics_match_restated_reference():
    g = torch.Generator().manual_seed(0)
    S, N, C = 100, 57, 2
    samples = torch.randn(S, N, C, generator=g) * 0.7 + torch.tensor([0.3, 0.6])
    label = torch.randint(0, C, (N,), generator=g)
    mv = stats.majority_voting_for_mc_samples(samples)
    assert torch.equal(mv, orc.majority_vote(samples))
    assert torch.equal(stats.majority_voting_for_mc_samples(list(samples)), mv)
    temp = 0.1737
    conf = stats.compute_ensemble_confidence(samples, temp)
    assert torch.allclose(conf, orc.ensemble_confidence(samples, temp), atol=1e-7)
    assert torch.allclose(stats.compute_ece(conf, label), orc.ece_l1(conf, label), atol=1e-7)
    for mine, ref in zip(stats.compute_mean_piws_for_class(samples, mv, label),
                         orc.mean_piw_per_class(samples, mv, label)):
        assert torch.allclose(mine, ref, equal_nan=True)
    for mine, ref in zip(stats.calculate_variances(samples, mv, label), orc.class_variances(samples, mv, label)):
        assert torch.allclose(mine, ref)
    assert float(stats.compute_accuracy(mv, label)) == pytest.approx(float((mv == label).float().mean()))


# ------------------------------------------------------------------ statistics vs the REFERENCE's own functions
def _stats_cases():
    from tests.golden_util import Fixture

    fx = Fixture("stats_reference")
    return fx, [c["tag"] for c in fx.meta["cases"]]


@pytest.mark.parametrize("tag", _stats_cases()[1])
def test_statistics_match_reference_functions(tag):
    """stats.py vs outputs of the reference's own majority_voting_for_mc_samples / compute_mean_piws_for_class /
    calculate_variances / convert_to_prob / compute_ensemble_confidence (classification_train_separately.py:51-68,
    :102-140, :143-174, :392-398, :425-447), lifted with ``ast`` and executed by tests/golden/make_golden.py::stats_fixture
    -- incl. vote ties, classes nobody predicts, a single chain.  (ECE is NOT here: torchmetrics is absent, see
    test_ece_known_answer_parity_unpinned.)"""
    fx, _ = _stats_cases()
    case = next(c for c in fx.meta["cases"] if c["tag"] == tag)
    samples, label = fx[f"{tag}/samples"], fx[f"{tag}/label"]
    mv = stats.majority_voting_for_mc_samples(samples)
    assert torch.equal(mv, fx[f"{tag}/mv"])
    assert torch.equal(stats.majority_voting_for_mc_samples([s for s in samples]), fx[f"{tag}/mv"])
    assert torch.equal(orc.majority_vote(samples), fx[f"{tag}/mv"])
    assert torch.allclose(stats.convert_to_prob(samples, case["temperature"]), fx[f"{tag}/prob"], atol=1e-7)
    assert torch.allclose(stats.compute_ensemble_confidence(samples, case["temperature"]), fx[f"{tag}/conf"], atol=1e-6)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # var() of a single chain warns (and is NaN) in the reference too
        piw = stats.compute_mean_piws_for_class(samples, mv, label)
        var = stats.calculate_variances(samples, mv, label)
    for mine, ref in zip(piw, (fx[f"{tag}/piw_correct"], fx[f"{tag}/piw_incorrect"])):
        assert mine.shape == ref.shape and torch.allclose(mine, ref, atol=1e-6, equal_nan=True), (tag, mine, ref)
    for mine, ref in zip(var, (fx[f"{tag}/var_correct"], fx[f"{tag}/var_incorrect"])):
        assert mine.shape == ref.shape and torch.allclose(mine, ref, atol=1e-6, equal_nan=True), (tag, mine, ref)
    # the oracle's restatements are pinned by the same fixture
    for mine, ref in zip(orc.mean_piw_per_class(samples, mv, label), (fx[f"{tag}/piw_correct"], fx[f"{tag}/piw_incorrect"])):
        assert torch.allclose(mine, ref, atol=1e-6, equal_nan=True)


def test_q_sample_y0_reparam_extract_match_reference():
    """diffusion_utils.q_sample / y_0_reparam / extract (a7, a8, a3) with per-row timesteps vs the reference's own
    outputs (tests/golden/aux_qsample_y0reparam.npz), plus the oracle's restatements."""
    from tests.golden_util import Fixture

    from nested_diffusion_b200 import diffusion_utils as du

    fx = Fixture("aux_qsample_y0reparam")
    m = fx.meta
    sd = {k[3:]: v for k, v in fx.arrays.items() if k.startswith("sd/")}
    cfg = argparse.Namespace(diffusion=argparse.Namespace(timesteps=m["T"]),
                             data=argparse.Namespace(num_classes=m["C"], dataset="ChestXRay"),
                             model=argparse.Namespace(data_dim=m["Dx"], arch="linear", feature_dim=m["F"], hidden_dim=m["H"]))
    model = nd.ConditionalModel(cfg, guidance=True)
    model.load_state_dict(sd)
    model.eval()
    t = fx["t"]
    y_t = du.q_sample(fx["y0"], fx["yhat"], fx["alphas_bar_sqrt"], fx["omabs"], t, noise=fx["noise"])
    assert torch.equal(y_t, fx["y_t"])
    torch.manual_seed(m["rng_seed"])
    assert torch.equal(du.q_sample(fx["y0"], fx["yhat"], fx["alphas_bar_sqrt"], fx["omabs"], t), fx["y_t_rng"])
    assert torch.equal(du.extract(fx["omabs"], t, y_t), fx["extract"])
    with torch.no_grad():
        y0r = du.y_0_reparam(model, fx["x"], fx["y_t"], fx["yhat"], fx["yhat"], t, fx["omabs"])
    assert torch.allclose(y0r, fx["y0_reparam"], rtol=1e-6, atol=1e-6)
    assert not y0r.requires_grad
    assert torch.equal(orc.q_sample(fx["y0"], fx["yhat"], fx["alphas_bar_sqrt"], fx["omabs"], t, fx["noise"]), fx["y_t"])
    with torch.no_grad():
        eps_fn = orc._Eps(sd, fx["x"], hoist=False)
        y0o = orc.y_0_reparam(eps_fn, fx["y_t"], fx["yhat"], fx["yhat"], t, fx["omabs"])
    assert torch.allclose(y0o, fx["y0_reparam"], rtol=1e-6, atol=1e-6)


def test_majority_vote_ties_go_to_smallest_label():
    # 4 chains, 2 vote class 1 and 2 vote class 0 -> the reference returns 0 (sorted unique + first argmax)
    s = torch.tensor([[[0.0, 1.0]], [[0.0, 1.0]], [[1.0, 0.0]], [[1.0, 0.0]]])
    assert int(stats.majority_voting_for_mc_samples(s)[0]) == 0 == int(orc.majority_vote(s)[0])
    s3 = torch.zeros(6, 1, 3)
    for i, c in enumerate([2, 2, 1, 1, 0, 2]):
        s3[i, 0, c] = 1
    assert int(stats.majority_voting_for_mc_samples(s3)[0]) == 2


def test_ece_known_answer_parity_unpinned():
    """ECE restated from torchmetrics 0.11.4 MulticlassCalibrationError(n_bins=10, norm='l1'); torchmetrics is not in
    the image, so this one statistic has only a hand-computed known answer: PARITY UNPINNED."""
    probs = torch.tensor([[0.9, 0.1], [0.8, 0.2], [0.4, 0.6], [0.55, 0.45]])
    target = torch.tensor([0, 1, 1, 0])
    # bins (0.5,0.6]: conf .55 acc 1 -> wait .6 falls in (0.5,0.6]: two items conf (.6,.55) acc (1,1);
    # (0.7,0.8]: conf .8 acc 0; (0.8,0.9]: conf .9 acc 1
    want = (abs(1 - 0.575) * 2 + abs(0 - 0.8) + abs(1 - 0.9)) / 4
    assert float(stats.compute_ece(probs, target)) == pytest.approx(want, abs=1e-6)
    assert float(orc.ece_l1(probs, target)) == pytest.approx(want, abs=1e-6)


# ------------------------------------------------------------------ sharding + gather (gloo, 2 ranks)
def test_shard_bounds_partition():
    for n in (1, 7, 70, 71, 1024):
        for w in (1, 2, 3, 8):
            spans = [nd.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_image_tiles_are_equal_and_cover_the_batch():
    """NestedEnsemble's rows-per-call cap (32 768 rows per member: longer launches lose L2 locality)."""
    from nested_diffusion_b200.ensemble import NestedEnsemble, image_tiles

    assert NestedEnsemble.MAX_ROWS_PER_CALL == 32768
    assert image_tiles(1024, 20, 32768) == [(0, 1024)]                              # config 3: one call
    assert image_tiles(16384, 10, 32768) == [(i * 2731, min(16384, (i + 1) * 2731)) for i in range(6)]
    assert image_tiles(1024, 1000, 32768) == [(i * 32, (i + 1) * 32) for i in range(32)]
    assert image_tiles(3, 100000, 32768) == [(0, 1), (1, 2), (2, 3)]                # draws beyond the cap: one image per call
    assert image_tiles(0, 20, 32768) == [(0, 0)]
    for n, d, cap in ((70, 20, 32768), (10, 3, 10), (977, 7, 500), (5, 1, 1)):
        tiles = image_tiles(n, d, cap)
        assert tiles[0][0] == 0 and tiles[-1][1] == n and all(a[1] == b[0] for a, b in zip(tiles, tiles[1:]))
        sizes = [b - a for a, b in tiles]
        assert max(sizes) * d <= max(cap, d) and max(sizes) - min(sizes) <= max(1, len(tiles) - 1)
        assert len(tiles) == -(-n // max(1, min(n, cap // d)))                      # the fewest tiles the cap allows


def test_weighted_bounds_partition():
    """Speed-weighted image tiles: contiguous, exhaustive, proportional; degenerate weights give empty tiles, not errors."""
    for n in (1, 7, 70, 1024):
        for w in ([1.0], [1.0, 1.0], [1.0, 0.9, 1.1], [0.96, 1.0, 1.0, 1.02, 0.99, 1.0, 0.95, 1.0], [1.0, 0.0, 1.0]):
            spans = nd.weighted_bounds(n, w)
            assert len(spans) == len(w) and spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] and spans[i][0] <= spans[i][1] for i in range(len(w) - 1))
            for (a, b), wi in zip(spans, w):
                assert abs((b - a) - n * wi / sum(w)) <= 1.0
    assert nd.weighted_bounds(1024, [1.0] * 8) == [nd.shard_bounds(1024, r, 8) for r in range(8)]
    with pytest.raises(ValueError):
        nd.weighted_bounds(10, [0.0, 0.0])


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import nested_diffusion_b200 as nd
from nested_diffusion_b200 import stats
rank, world, n = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[5])
dist.init_process_group("gloo", rank=rank, world_size=world)
g = torch.Generator().manual_seed(5)
full = torch.randn(n, 100, 2, generator=g)           # [N, K*D, C] "samples" every rank could have produced
lo, hi = nd.shard_bounds(n, rank, world)
got = nd.gather_image_shards(full[lo:hi].contiguous(), n)
assert torch.equal(got, full), "gathered tensor differs from the unsharded one"
# speed-weighted (uneven) image tiles gather to the same tensor
bounds = nd.weighted_bounds(n, [1.0, 0.55])
wlo, whi = bounds[rank]
got_w = nd.gather_image_shards(full[wlo:whi].contiguous(), n, bounds=bounds)
assert torch.equal(got_w, full), "weighted shards gather differently"
mv = stats.majority_voting_for_mc_samples(got.permute(1, 0, 2))
ref = stats.majority_voting_for_mc_samples(full.permute(1, 0, 2))
assert torch.equal(mv, ref)
# the whole sharded entry point (host logic of config 4) with a stub ensemble whose "samples" are a pure function of the
# GLOBAL (member, draw, image) ids -- as the Philox streams are: any partition must gather to the single-process result
from nested_diffusion_b200.ensemble import EnsembleResult, NestedEnsemble, image_tiles

class Stub(NestedEnsemble):
    def __init__(self, K):
        self.members, self.models, self.member_ids = [None] * K, [], list(range(K))
        self.device, self.precision, self.max_rows_per_call = torch.device("cpu"), "auto", 64
        self.calls = []
    @property
    def K(self):
        return len(self.members)
    def sample(self, x, y0hats, draws, n_steps, alphas, omabs, *, seed=None, temperature=None, image_offset=0,
               images_total=0, **kw):
        self.calls.append((image_offset, x.shape[0], images_total))
        K, N, C = len(self.members), x.shape[0], y0hats.shape[-1]
        k = torch.arange(K).view(K, 1, 1, 1); d = torch.arange(draws).view(1, draws, 1, 1)
        i = (torch.arange(N) + image_offset).view(1, 1, N, 1); c = torch.arange(C).view(1, 1, 1, C)
        y = (seed + 1000.0 * k + 10.0 * d + 0.001 * i + 0.0001 * c + x[:, :1].view(1, 1, N, 1) + y0hats.unsqueeze(1)).float()
        return EnsembleResult(y, torch.softmax(-y / temperature, -1) if temperature is not None else None,
                              (image_offset, image_offset + N))

K, D, C = 3, 4, 2
x = torch.randn(n, 5, generator=g); yh = torch.softmax(torch.randn(K, n, C, generator=g), -1)
ens = Stub(K)
y0, probs = nd.sample_ensemble(ens, x, yh, D, 10, None, None, seed=7, temperature=0.5)
lo, hi = nd.shard_bounds(n, rank, world)
assert ens.calls == [(lo, hi - lo, n)], ens.calls                       # this rank sampled exactly its image tile
single = Stub(K).sample(x, yh, D, 10, None, None, seed=7, temperature=0.5)
want_y = single.y0.permute(2, 0, 1, 3).reshape(n, K * D, C)
assert tuple(y0.shape) == (n, K * D, C) and torch.equal(y0, want_y)
assert torch.equal(probs, single.probs.permute(2, 0, 1, 3).reshape(n, K * D, C))
yw, _ = nd.sample_ensemble(Stub(K), x, yh, D, 10, None, None, seed=7, temperature=0.5, shard_weights=[1.0, 0.4])
assert torch.equal(yw, want_y), "speed-weighted tiles must not change the gathered samples"
y_none, p_none = nd.sample_ensemble(Stub(K), x, yh, D, 10, None, None, seed=7)
assert torch.equal(y_none, want_y) and p_none is None
# without a seed the ranks must agree on one (rank 0 draws it, broadcast): every rank ends with the same tensor
torch.manual_seed(100 + rank)                                            # different local generators on purpose
y_auto, _ = nd.sample_ensemble(Stub(K), x, yh, D, 10, None, None)
both = [torch.empty_like(y_auto) for _ in range(world)]
dist.all_gather(both, y_auto)
assert all(torch.equal(b, both[0]) for b in both), "ranks sampled with different seeds"
try:
    nd.sample_ensemble(Stub(K), x, yh, D, 10, None, None, seed=7, shard_weights=[1.0])
    raise SystemExit("shard_weights of the wrong length must raise")
except ValueError:
    pass
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


@pytest.mark.parametrize("n_images", [70, 7])
def test_gather_image_shards_two_ranks_gloo(n_images, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + (os.getpid() + n_images) % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", str(n_images), port],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


# ------------------------------------------------------------------ Philox oracle
def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        got = pho.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(v) for v in got] == want


def test_philox_noise_is_standard_normal_and_partition_invariant():
    z = pho.noise_tensor(1234, K=2, D=3, S=50, N=40, C=2)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    # kurtosis of a normal is 3
    assert abs(((z - z.mean()) ** 4).mean() / z.var() ** 2 - 3) < 0.15
    part = pho.noise_tensor(1234, K=2, D=3, S=50, N=15, C=2, image_offset=25, images_total=40)
    assert np.array_equal(part, z[:, :, :, 25:40])
    member1 = pho.noise_tensor(1234, K=1, D=3, S=50, N=40, C=2, member_ids=[1])
    assert np.array_equal(member1[0], z[1])


# ------------------------------------------------------------------ runner shim (host logic only)
def test_runner_metrics_and_checkpoint_loader(tmp_path):
    from nested_diffusion_b200.runner import (SampleCache, ensemble_metrics, load_noise_estimators,
                                              temperature_for)

    g = torch.Generator().manual_seed(3)
    cache = SampleCache()
    for _ in range(3):
        cache.y0.append(torch.randn(40, 9, 2, generator=g) * 0.6 + torch.tensor([0.2, 0.7]))
        cache.target.append(torch.randint(0, 2, (9,), generator=g))
    m = ensemble_metrics(cache, 0.1737)
    allv = torch.cat(cache.y0, dim=1)
    tgt = torch.cat(cache.target)
    mv = torch.cat([orc.majority_vote(s) for s in cache.y0])
    assert torch.equal(m["majority_vote"], mv)
    assert float(m["accuracy"]) == pytest.approx(float((mv == tgt).float().mean()))
    prob = torch.cat([orc.ensemble_confidence(s, 0.1737) for s in cache.y0])
    assert float(m["ece"]) == pytest.approx(float(orc.ece_l1(prob, tgt)), abs=1e-6)
    for mine, ref in zip((m["var_correct"], m["var_incorrect"]), orc.class_variances(allv, mv, tgt)):
        assert torch.allclose(mine, ref)
    assert temperature_for("ISICSkinCancerAtkPGD") == 0.3162
    with pytest.raises(NotImplementedError):
        temperature_for("MNIST")
    cfg = argparse.Namespace(diffusion=argparse.Namespace(timesteps=5, include_guidance=True),
                             data=argparse.Namespace(num_classes=2, dataset="ChestXRay"),
                             model=argparse.Namespace(data_dim=10, arch="linear", feature_dim=8, hidden_dim=6))
    src = nd.ConditionalModel(cfg, guidance=True)
    path = tmp_path / "diffu0_ckpt_best.pth"
    torch.save({"noise_estimator": src.state_dict(), "optimizer": {}, "epoch": 7}, path)  # the reference's layout
    (loaded,) = load_noise_estimators(cfg, [str(path)], "cpu")
    assert not loaded.training
    for k, v in src.state_dict().items():
        assert torch.equal(v, loaded.state_dict()[k])


def test_packed_cache_key_follows_content_not_name(tmp_path):
    """the on-disk packed cache is keyed by the checkpoint's CONTENT (+ precision + ABI version)."""
    from nested_diffusion_b200.runner import _content_key, _read_image, _write_image

    a, b, c = tmp_path / "a.pth", tmp_path / "b.pth", tmp_path / "c.pth"
    a.write_bytes(b"x" * 100000)
    b.write_bytes(b"x" * 100000)
    c.write_bytes(b"x" * 99999 + b"y")
    assert _content_key(str(a), "fp16") == _content_key(str(b), "fp16")
    assert _content_key(str(a), "fp16") != _content_key(str(c), "fp16")
    assert _content_key(str(a), "fp16") != _content_key(str(a), "fp32x")
    assert _content_key(str(a), "auto").endswith(f"-auto-abi{_capi.load().ladine_version()}")
    img = torch.arange(1000, dtype=torch.int64).view(torch.uint8)
    _write_image(str(tmp_path / "i.ladine"), img)
    assert torch.equal(_read_image(str(tmp_path / "i.ladine")), img)
    assert [p.name for p in tmp_path.iterdir() if ".tmp" in p.name] == []


def test_cache_validity_rules_on_cpu_tensors():
    """engine._Cached (the rule behind the pack / encoder / feature caches) and diffusion_utils._tensor_key (draws-ahead):
    an in-place edit invalidates, a storage swap with the same values (the runner's CPU<->GPU shuttle) does not, a
    storage swap with new values does."""
    from nested_diffusion_b200 import diffusion_utils as du
    from nested_diffusion_b200 import engine

    lin = torch.nn.Linear(64, 32)
    tensors = lambda: [(k, v) for k, v in lin.state_dict(keep_vars=True).items()]
    c = engine._Cached(tensors(), "fp16", "packed")
    assert c.valid_for(tensors(), "fp16") and not c.valid_for(tensors(), "bf16")
    lin.weight.data = lin.weight.data.clone()                      # new storage, same values
    assert c.valid_for(tensors(), "fp16")
    lin.bias.data = lin.bias.data + 1.0                            # new storage, new values, no version bump
    assert not c.valid_for(tensors(), "fp16")
    c = engine._Cached(tensors(), "fp16", "packed")
    with torch.no_grad():
        lin.weight.mul_(1.5)                                       # in-place: version counter
    assert not c.valid_for(tensors(), "fp16")

    x = torch.randn(6, 4)
    k0 = du._tensor_key(x)
    assert du._tensor_key(x) == k0 and du._tensor_key(x[:]) == k0  # a fresh view of the same data is the same input
    assert du._tensor_key(x[1:]) != k0 and du._tensor_key(x.clone()) != k0
    x.add_(1.0)
    assert du._tensor_key(x) != k0                                 # in-place change
    assert du.set_draws_ahead(20) == 0 and du.set_draws_ahead(0) == 20


def test_coef_table_is_remembered_until_the_schedule_changes():
    """schedule.coef_table caches per (alphas, one_minus_alphas_bar_sqrt, n_steps): no device->host copy (= stream
    synchronisation) on every one of the runner's K x 20 calls; an in-place edit or another tensor recomputes."""
    alphas, omabs = schedule.schedule_tensors(schedule.make_beta_schedule("linear", 50, 1e-4, 0.02))
    t1 = schedule.coef_table(alphas, omabs, 50)
    assert schedule.coef_table(alphas, omabs, 50) is t1
    assert schedule.coef_table(alphas, omabs, 40) is not t1 and schedule.coef_table(alphas, omabs, 40).shape[0] == 40
    assert schedule.coef_table(alphas, omabs, 50) is t1                       # still remembered
    a2 = alphas.clone()
    t2 = schedule.coef_table(a2, omabs, 50)
    assert t2 is not t1 and torch.equal(t2, t1)
    a2.mul_(0.999)                                                            # in-place: version counter
    t3 = schedule.coef_table(a2, omabs, 50)
    assert t3 is not t2 and not torch.equal(t3, t2)
    assert torch.equal(t3, schedule._coef_table(a2, omabs, 50))
    for i in range(20):                                                       # bounded: old entries fall out
        schedule.coef_table(alphas.clone(), omabs, 50)
    assert len(schedule._COEF_CACHE) <= schedule._COEF_CACHE_SIZE


def test_partitions_property_based():
    """shard_bounds / weighted_bounds / image_tiles on random sizes (hypothesis): contiguous, exhaustive, within their caps."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from nested_diffusion_b200.ensemble import image_tiles

    @settings(max_examples=200, deadline=None)
    @given(n=st.integers(0, 5000), world=st.integers(1, 16))
    def shards(n, world):
        spans = [nd.shard_bounds(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
        assert max(sizes) == nd.ensemble.padded_shard_size(n, world)

    @settings(max_examples=200, deadline=None)
    @given(n=st.integers(0, 5000), w=st.lists(st.floats(0.05, 4.0), min_size=1, max_size=16))
    def weighted(n, w):
        spans = nd.weighted_bounds(n, w)
        assert len(spans) == len(w) and spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(spans, spans[1:]))
        assert all(abs((b - a) - n * wi / sum(w)) <= 1.0 for (a, b), wi in zip(spans, w))

    @settings(max_examples=300, deadline=None)
    @given(n=st.integers(0, 20000), d=st.integers(1, 2000), cap=st.integers(1, 70000))
    def tiles(n, d, cap):
        t = image_tiles(n, d, cap)
        assert t[0][0] == 0 and t[-1][1] == n and all(a[1] == b[0] for a, b in zip(t, t[1:]))
        if n:
            per = max(1, min(n, cap // d))
            sizes = [b - a for a, b in t]
            assert min(sizes) >= 1 and max(sizes) <= per and len(t) == -(-n // per)
            assert max(sizes) - min(sizes) <= len(t)        # (nearly) equal tiles: no small remainder call

    shards(); weighted(); tiles()


def test_packed_image_header_checks_without_a_device():
    """ladine_image_info (host only): an image built to the documented layout is accepted; a flipped payload byte, a
    truncated buffer, another layout version, a foreign magic and a too-short buffer are refused with a reason."""
    import struct

    lib = _capi.load()

    def image(magic=b"LADINEM", layout=1, abi=None, payload=None, dims=(256, 256, 2, 2, 200, 1, 2, 0)):
        payload = np.arange(4096, dtype=np.uint8).tobytes() * 3 if payload is None else payload
        buf = np.zeros(128 + len(payload), dtype=np.uint8)
        buf[128:] = np.frombuffer(payload, dtype=np.uint8)
        ck = lib.ladine_image_checksum(buf[128:].ctypes.data, len(payload))
        hdr = struct.pack("<8sII20iQQQ", magic, lib.ladine_version() if abi is None else abi, layout,
                          *(list(dims) + [0] * (20 - len(dims))), len(payload), ck, 0)
        assert len(hdr) == 120
        buf[:120] = np.frombuffer(hdr, dtype=np.uint8)
        return buf

    def info(buf, n=None):
        kind, dims, why = ctypes.c_int32(), (ctypes.c_int32 * 20)(), ctypes.c_char_p()
        rc = lib.ladine_image_info(buf.ctypes.data, len(buf) if n is None else n, ctypes.byref(kind), dims, ctypes.byref(why))
        return rc, kind.value, list(dims), (why.value or b"").decode()

    rc, kind, dims, why = info(image())
    assert (rc, kind, dims[:8], why) == (0, 1, [256, 256, 2, 2, 200, 1, 2, 0], "")
    rc, kind, dims, _ = info(image(magic=b"LADINEE", dims=(300, 128, 256, 0)))
    assert (rc, kind, dims[:3]) == (0, 2, [300, 128, 256])
    bad = image()
    bad[128 + 1000] ^= 0x10
    assert info(bad)[0] < 0 and "checksum" in info(bad)[3]
    assert info(image(), n=128 + 4096)[0] < 0 and "truncated" in info(image(), n=128 + 4096)[3]
    assert "different library version" in info(image(layout=2))[3]
    assert "different library version" in info(image(abi=lib.ladine_version() + 1))[3]
    assert info(image(magic=b"NOTLADN"))[0] < 0 and "magic" in info(image(magic=b"NOTLADN"))[3]
    assert info(image(), n=64)[0] < 0 and "shorter" in info(image(), n=64)[3]
    # the checksum depends on every byte and on the length
    a = np.arange(1000, dtype=np.uint8)
    c0 = lib.ladine_image_checksum(a.ctypes.data, 1000)
    b = a.copy(); b[999] ^= 1
    assert lib.ladine_image_checksum(b.ctypes.data, 1000) != c0 and lib.ladine_image_checksum(a.ctypes.data, 999) != c0
