"""CPU-only check of the PRODUCT's MLP device code and tree planner: k_mlp.cuh and
mlp_plan.cu are compiled for the host (tests/hostemu, single-lane group) and compared
with the oracle on identical injected tapes.  This validates, without a GPU, the
consumption-ordinal arithmetic ("bit-exact indexing"), the tree cut at any threshold,
the bottom-up combine, the kinds, and the per-coordinate reflection."""
import ctypes as C

import numpy as np
import pytest

from helpers import emu_solve, oracle_solve, random_tapes, rel_err, zoo

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def _names():
    import hdpde_b200 as H
    return list(zoo(H).keys())


@pytest.mark.parametrize("name", _names())
def test_emulated_device_code_matches_oracle(H, O, emu, name):
    prob, alg, dt = zoo(H)[name]
    tapes = random_tapes(H, prob, alg, dt, seed=11)
    uo, info = oracle_solve(H, O, prob, alg, dt, tapes)
    assert info["rc"] == 0 and not info["tape_pos"]["overrun"]
    # the oracle consumed exactly the closed-form counts
    assert info["tape_pos"]["r"] == tapes[0].size and info["tape_pos"]["n"] == tapes[1].size
    seen_units = set()
    for target in (0, 8, 64, 1000, 100000):
        u, rc, st = emu_solve(H, emu, prob, alg, dt, tapes=tapes, target_units=target)
        assert rc == 0
        assert rel_err(float(u[0]), uo) <= TOL[dt], (name, target, u[0], uo)
        # reflection decisions are exact: identical number of wall hits
        assert st.bounces == info["bounces"], (name, target)
        seen_units.add(int(st.n_units))
        if target == 0:
            assert st.n_units == 1 and float(u[0]) == uo  # no cut, same summation order: bit-equal
    assert len(seen_units) >= 3  # the thresholds really produced different cuts


def test_cut_plans_cover_the_tree(H, emu):
    """Every plan consumes exactly the closed-form number of ordinals (no gaps, no overlap):
    with a tape of the exact length any mis-indexing would read zeros or overrun."""
    import ctypes as C
    out = (C.c_int64 * 8)()
    for (M, L, K, nl) in ((4, 4, 10, 0), (4, 5, 10, 0), (3, 4, 3, 1), (2, 6, 2, 1)):
        c = H.mlp_counts(M, L, K, bool(nl))
        for thresh in (1.0, 24.0, 100.0, 1e4, 1e9):
            emu.hdp_emu_plan_summary(M, L, K, nl, C.c_double(thresh), 0, out)
            assert out[5] == c["path_steps"] and out[6] == c["time_uniforms"] and out[7] == c["sampler_calls"]
            assert out[0] >= 1 and out[2] >= 1


@pytest.mark.parametrize("nt", [1, 3, 8])
def test_compat_mlt_partition(H, O, emu, nt):
    """`_ml_picard_mlt` + `_get_loop_num` (src/MLP.jl:218-404, quirk Q4) reproduced behind
    compat_mlt_threads, on one shared tape consumed task by task."""
    prob, alg, dt = zoo(H)["allen_cahn_neumann_d5"]
    tapes = random_tapes(H, prob, alg, dt, seed=5, compat=nt)
    pb, keep = prob.to_c()
    sm, k2 = alg.mc_sample._descriptor(prob.d)
    rng = O.Rng.tape(dt, *tapes)
    uo, info = O.ml_picard_mlt(pb, sm, alg.M, alg.L, alg.K, prob.x, nt, [rng] * (nt + 1), parallel=False)
    pos = rng.tape_pos()
    assert info["rc"] == 0 and not pos["overrun"] and pos["r"] == tapes[0].size and pos["n"] == tapes[1].size
    for target in (0, 64, 5000):
        u, rc, st = emu_solve(H, emu, prob, alg, dt, tapes=tapes, target_units=target, compat=nt)
        assert rc == 0 and rel_err(float(u[0]), uo) <= 1e-12
        assert st.bounces == info["bounces"]


def test_philox_mode_matches_oracle_philox(H, O, emu):
    """RNG mode is tape mode with the tape generated on the fly: the emulated generator and
    the oracle's Philox-by-ordinal stream address the same draws."""
    prob, alg, dt = zoo(H)["nonlocal_ac_neumann_d5"]
    pb, keep = prob.to_c()
    sm, k2 = alg.mc_sample._descriptor(prob.d)
    for seed in (0, 77):
        uo, info = O.ml_picard(pb, sm, alg.M, alg.L, alg.K, prob.x, O.Rng.philox(seed, dt, False))
        for target in (0, 1000):
            u, rc, st = emu_solve(H, emu, prob, alg, dt, seed=seed, target_units=target)
            assert rc == 0 and rel_err(float(u[0]), uo) <= 1e-11


def test_batch_of_solves_uses_consecutive_seeds(H, O, emu):
    prob, alg, dt = zoo(H)["heat_neumann_d5"]
    pb, keep = prob.to_c()
    sm, k2 = alg.mc_sample._descriptor(prob.d)
    u, rc, st = emu_solve(H, emu, prob, alg, dt, seed=40, n_solves=3, target_units=50)
    for i in range(3):
        uo, _ = O.ml_picard(pb, sm, alg.M, alg.L, alg.K, prob.x, O.Rng.philox(40 + i, dt, False))
        assert rel_err(float(u[i]), uo) <= 1e-11


def test_start_outside_domain_is_reported(H, O, emu):
    """Q5: NormalSampling + neumann_bc can leave the box; the reference errors in `_reflect`."""
    d = 3
    prob = H.PIDEProblem(0.0, 0.1, np.zeros(d), (0.0, 0.5), H.ExpSumSq(1.0), H.FNLFisher(),
                         neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    alg = H.MLP(M=2, L=3, K=2, mc_sample=H.NormalSampling(5.0))
    tapes = random_tapes(H, prob, alg, np.float64, seed=2)
    uo, info = oracle_solve(H, O, prob, alg, np.float64, tapes)
    assert info["rc"] == H.capi.HDP_ERR_START_OUTSIDE_DOMAIN
    u, rc, st = emu_solve(H, emu, prob, alg, np.float64, tapes=tapes)
    assert rc == H.capi.HDP_ERR_START_OUTSIDE_DOMAIN


def test_product_generator_matches_its_definition(O, emu):
    """The product's table-driven fp64 Box–Muller (hdp_device.cuh, compiled for the host) against the libm
    evaluation of the same definition in the oracle: uniforms bit-equal, normals to a few ulp, over every
    stream and over the whole range of the radius (including the entries of the log table next to u1 = 1)."""
    import ctypes as C
    emu.hdp_emu_draw_f64.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_double)]
    out = (C.c_double * 2)()
    worst = 0.0
    for stream in range(6):
        for ord_ in range(0, 3000, 1 if stream == 1 else 37):
            grp = (ord_ * 7) % 50
            emu.hdp_emu_draw_f64(99, stream, ord_, grp, 0, out)
            assert out[0] == O.draw(np.float64, 99, stream, ord_, 2 * grp, False)
            assert out[1] == O.draw(np.float64, 99, stream, ord_, 2 * grp + 1, False)
            emu.hdp_emu_draw_f64(99, stream, ord_, grp, 1, out)
            a = O.draw(np.float64, 99, stream, ord_, 2 * grp, True)
            b = O.draw(np.float64, 99, stream, ord_, 2 * grp + 1, True)
            worst = max(worst, abs(out[0] - a) / max(1.0, abs(a)), abs(out[1] - b) / max(1.0, abs(b)))
    assert worst <= 2e-15, worst


# ---------------------------------------------------------------------------------------------
# one solve sharded over several ranks (MlpSharding): the emulation runs the ranks one after the
# other, sums their exchange buffers as the all-reduce would and lets every rank finish the crown
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["allen_cahn_neumann_d5", "nonlocal_ac_neumann_d5", "replicator_d2", "black_scholes_d20",
                                  "shifted_normal_d3_L5", "heat_neumann_f32_d3"])
def test_sharded_solve_matches_oracle_for_any_rank_count(H, O, emu, name):
    prob, alg, dt = zoo(H)[name]
    tapes = random_tapes(H, prob, alg, dt, seed=13)
    uo, info = oracle_solve(H, O, prob, alg, dt, tapes)
    for target in (0, 40, 800, 100000):
        u1, rc, st1 = emu_solve(H, emu, prob, alg, dt, tapes=tapes, target_units=target)
        assert rc == 0
        for nranks in (2, 3, 8):
            u, rc, st = emu_solve(H, emu, prob, alg, dt, tapes=tapes, target_units=target, nranks=nranks)
            assert rc == 0, (name, target, nranks)  # (also: every rank finished with the same bits)
            assert rel_err(float(u[0]), uo) <= TOL[dt], (name, target, nranks, u[0], uo)
            assert abs(float(u[0]) - float(u1[0])) <= 4 * TOL[dt] * 1e-3 * max(1.0, abs(uo))  # summation order only
            assert st.bounces == info["bounces"], (name, target, nranks)  # every wall hit counted once
            assert st.exchange_bytes > 0


def test_sharded_batch_and_compat_partition(H, O, emu):
    prob, alg, dt = zoo(H)["heat_neumann_d5"]
    pb, keep = prob.to_c()
    sm, k2 = alg.mc_sample._descriptor(prob.d)
    u, rc, st = emu_solve(H, emu, prob, alg, dt, seed=40, n_solves=3, target_units=300, nranks=4)
    assert rc == 0
    for i in range(3):
        uo, _ = O.ml_picard(pb, sm, alg.M, alg.L, alg.K, prob.x, O.Rng.philox(40 + i, dt, False))
        assert rel_err(float(u[i]), uo) <= 1e-11
    # the threaded variant's partition of the root call, sharded
    prob, alg, dt = zoo(H)["allen_cahn_neumann_d5"]
    nt = 3
    tapes = random_tapes(H, prob, alg, dt, seed=5, compat=nt)
    pb, keep = prob.to_c()
    sm, k2 = alg.mc_sample._descriptor(prob.d)
    rng = O.Rng.tape(dt, *tapes)
    uo, info = O.ml_picard_mlt(pb, sm, alg.M, alg.L, alg.K, prob.x, nt, [rng] * (nt + 1), parallel=False)
    u, rc, st = emu_solve(H, emu, prob, alg, dt, tapes=tapes, target_units=64, compat=nt, nranks=2)
    assert rc == 0 and rel_err(float(u[0]), uo) <= 1e-12 and st.bounces == info["bounces"]


def test_sharding_exchanges_one_double_per_crown_node(H, emu):
    """BASELINE C4 (M=4, L=5, K=10, non-local) cut as finely as 8 GPUs need: hundreds of thousands of units, a few
    thousand crown nodes — the all-reduce carries the latter — and a balanced split of the modelled work."""
    import ctypes as C
    out = (C.c_int64 * 8)()
    plan = (C.c_int64 * 8)()
    for thresh in (400.0, 90.0, 45.0):   # (45: the level-1 sites explicit, 5.6 M units)
        emu.hdp_emu_plan_summary(4, 5, 10, 1, C.c_double(thresh), 0, plan)
        n_units = plan[2]
        units = 0
        for rank in range(8):
            emu.hdp_emu_sharding_summary(4, 5, 10, 1, C.c_double(thresh), 8, rank, out)
            units += out[6]
            assert out[0] == out[1] and out[0] < n_units // 20, (thresh, list(out), n_units)
            assert out[7] <= 1020, (thresh, list(out))  # busiest rank within 2 % of the mean
        assert units == n_units  # every unit has exactly one owner


# ---------------------------------------------------------------------------------------------
# Monte-Carlo error bar (cfg.std_err_out)
# ---------------------------------------------------------------------------------------------
def test_error_bar_is_the_level_statistics_of_the_root(H, O, emu):
    """Heat equation (f = 0): only the g-term contributes; its sample statistics can be recomputed from the tape."""
    prob, alg, dt = zoo(H)["heat_d10"]
    tapes = random_tapes(H, prob, alg, dt, seed=3)
    cn = H.mlp_counts(alg.M, alg.L, alg.K)
    ng = alg.M ** alg.L
    # the g-term steps are the last M^L path-steps of the root call: x2 = x + sigma*sqrt(T)*xi, g = sum(x2^2)
    xi = tapes[1][cn["path_steps"] - ng:]
    g = np.sum((prob.x[None, :] + 0.1 * np.sqrt(0.5) * xi) ** 2, axis=1)
    se_want = np.sqrt(g.var(ddof=1) / ng)
    ref = None
    for target, nranks in ((0, 1), (30, 1), (3000, 1), (30, 2), (3000, 8)):
        u, rc, st = emu_solve(H, emu, prob, alg, dt, tapes=tapes, target_units=target, nranks=nranks, std_err=True)
        assert rc == 0
        assert st.level_n[alg.L] == ng and abs(st.level_sum[alg.L] - g.sum()) <= 1e-9 * g.sum()
        assert abs(st.level_sumsq[alg.L] - np.sum(g * g)) <= 1e-9 * np.sum(g * g)
        assert [st.level_n[l] for l in range(alg.L)] == [alg.M ** (alg.L - l) for l in range(alg.L)]
        assert abs(st.std_error - se_want) <= 1e-6 * se_want and abs(float(st.se[0]) - se_want) <= 1e-6 * se_want
        ref = ref or st.std_error


def test_error_bar_covers_the_scatter_of_independent_solves(H, emu):
    """Allen–Cahn with reflection: the reported standard error matches the empirical spread over seeds."""
    prob, alg, dt = zoo(H)["allen_cahn_neumann_d5"]
    u, rc, st = emu_solve(H, emu, prob, alg, dt, seed=7, n_solves=48, target_units=48 * 40, std_err=True)
    assert rc == 0
    sd = float(np.std(u, ddof=1))
    se = float(np.sqrt(np.mean(st.se ** 2)))
    assert 0.6 * sd <= se <= 1.6 * sd, (sd, se)


# ---------------------------------------------------------------------------------------------
# the specialised coordinate loops in RNG mode with an even d (no partial vec-group), on the CPU
# ---------------------------------------------------------------------------------------------
def _leaf_cases(H):
    c = {}
    c["sym_d6"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(6), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                         neumann_bc=(-0.5, 0.5)), H.MLP(M=3, L=3, K=2), np.float64)     # partial last batches (M = 3)
    c["sym_d4_M4"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(4), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                            neumann_bc=(-0.5, 0.5)), H.MLP(M=4, L=3, K=3), np.float64)
    c["box_mu_d4"] = (H.ParabolicPDEProblem(0.3, 0.8, np.full(4, 0.1), (0.0, 0.5), g=H.SumSq(1.0), f=H.FLinear(0.3),
                                            neumann_bc=(-0.3, 0.7)), H.MLP(M=3, L=3, K=2), np.float64)  # mu != 0, general scalar box
    c["nobc_d8"] = (H.ParabolicPDEProblem(0.05, 0.3, np.zeros(8), (0.0, 0.5), g=H.InvSumSq(2.0, 0.4), f=H.FAllenCahn(-1.0)),
                    H.MLP(M=5, L=2, K=2), np.float64)
    c["wild_sym_d2"] = (H.ParabolicPDEProblem(0.0, 30.0, np.zeros(2), (0.0, 1.0), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                              neumann_bc=(-0.5, 0.5)), H.MLP(M=3, L=3, K=2), np.float64)  # many double hits, far excursions
    c["nonlocal_dead_d4"] = (H.PIDEProblem(0.0, 0.6, np.zeros(4), (0.0, 0.5), H.ExpSumSq(0.25), H.FNLAllenCahnDiff(),
                                           neumann_bc=(-0.5, 0.5)),
                             H.MLP(M=3, L=3, K=3, mc_sample=H.UniformSampling(-0.5, 0.5)), np.float64)  # level-0 samples are dead draws
    c["replicator_live_d2"] = (H.PIDEProblem(0.0, 0.1, np.zeros(2), (0.0, 0.5), H.GaussDensity(1e-2), H.FNLReplicatorMutator(1.0)),
                               H.MLP(M=3, L=3, K=2, mc_sample=H.NormalSampling(1.0)), np.float64)       # f reads z: sites not batched, g-term is
    c["sym_d8_f32"] = (H.ParabolicPDEProblem(0.0, 1.0, np.zeros(8, dtype=np.float32), (0.0, 0.5), g=H.ExpSumSq(0.25),
                                             f=H.FAllenCahn(1.0), neumann_bc=(-0.5, 0.5)), H.MLP(M=3, L=3, K=2), np.float32)
    return c


@pytest.mark.parametrize("name", ["sym_d6", "sym_d4_M4", "box_mu_d4", "nobc_d8", "wild_sym_d2", "nonlocal_dead_d4",
                                  "replicator_live_d2", "sym_d8_f32"])
def test_fast_loops_match_the_oracle_on_the_same_philox_draws(H, O, emu, name):
    """RNG mode takes the specialised loops (no boundary / symmetric box / scalar box, mu = 0 or not, dead sampler draws);
    the oracle draws the same Philox numbers by ordinal.  Values to the generator's accuracy (libm vs the product's
    tables), wall hits exact, for every cut and for a sharded run.  (These cases were written for the leaf-batch
    experiment of profiles/r2p_leaf_batch_experiment.md, which they passed; it was slower and is not in the product.)"""
    prob, alg, dt = _leaf_cases(H)[name]
    pb, keep = prob.to_c()
    sm, k2 = alg.mc_sample._descriptor(prob.d)
    normal = isinstance(alg.mc_sample, H.NormalSampling)
    for seed in (3, 1234567):
        uo, info = O.ml_picard(pb, sm, alg.M, alg.L, alg.K, np.asarray(prob.x, dtype=dt), O.Rng.philox(seed, dt, normal))
        assert info["rc"] == 0
        tol = 1e-11 if dt == np.float64 else 2e-5
        for target, nranks in ((0, 1), (7, 1), (200, 1), (100000, 1), (200, 3)):
            u, rc, st = emu_solve(H, emu, prob, alg, dt, seed=seed, target_units=target, nranks=nranks)
            assert rc == 0
            assert rel_err(float(u[0]), uo) <= tol, (name, seed, target, nranks, float(u[0]), uo)
            if dt == np.float64 and name != "wild_sym_d2":
                assert st.bounces == info["bounces"], (name, seed, target, nranks)


# ---------------------------------------------------------------------------------------------
# long non-local sites made explicit (whole-call children), ownership roots with explicit sites below them
# ---------------------------------------------------------------------------------------------
def test_long_sites_and_nested_ownership_match_the_oracle(H, O, emu):
    """thresh between W[1] and Ws[1] / 6: a level-1 site of a non-local problem (1 + K child calls) is far longer than a
    unit although each child call fits one — the planner makes such sites explicit, their children whole-call units.
    Sharded, with few ownership roots per rank asked for, the level-2 sites become the roots and the explicit level-1
    sites nest below them: geometry and combine of the owner run depth by depth.  Same tapes, every cut and rank count:
    the oracle's value, and its wall-hit count."""
    prob = H.PIDEProblem(0.0, 0.4, np.zeros(2), (0.0, 0.5), H.ExpSumSq(0.25), H.FNLFisher(), neumann_bc=(-0.5, 0.5))
    alg = H.MLP(M=3, L=3, K=8, mc_sample=H.UniformSampling(-0.5, 0.5))
    dt = np.float64
    tapes = random_tapes(H, prob, alg, dt, seed=5)
    uo, info = oracle_solve(H, O, prob, alg, dt, tapes)
    plan = (C.c_int64 * 8)()
    # W[1] = 18, Ws[1] = 167 (M = 3, K = 8): a 29-unit cut target keeps whole U_1 calls and splits the sites
    emu.hdp_emu_plan_summary(3, 3, 8, 1, C.c_double(20.0), 0, plan)
    n_units, n_sites = plan[2], plan[1]
    assert n_sites >= 9 * 3 and n_units > 5 * n_sites, list(plan)   # level-1 sites explicit, 1 + K whole calls each
    target = int(alg_work(H, alg) / 20.0)
    try:
        for own_items, nranks in ((0, 1), (0, 2), (0, 3), (1.0, 2), (0.5, 3), (2.0, 8)):
            emu.hdp_emu_set_own_items(C.c_double(own_items))
            u, rc, st = emu_solve(H, emu, prob, alg, dt, tapes=tapes, target_units=target, nranks=nranks)
            assert rc == 0
            assert rel_err(float(u[0]), uo) <= 1e-12, (own_items, nranks, float(u[0]), uo)
            assert st.bounces == info["bounces"], (own_items, nranks)
            assert st.n_sites == n_sites
    finally:
        emu.hdp_emu_set_own_items(C.c_double(0.0))


def test_crown_share_table_tiles_every_heavy_node(emu):
    """mlp_build_crown_jobs: jobs (k, 0..nb-1) in order, leaves and light nodes without jobs, at most 64 shares per
    node, and the shares — cut as prereduce_share cuts them — cover every unit and every site of a heavy node exactly
    once.  Trees of the strong-scaling cases (Allen-Cahn M=8, C4) at the library's thresholds and small ones forced."""
    emu.hdp_emu_crown_jobs_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_long, C.c_long,
                                             C.POINTER(C.c_int64)]
    out = (C.c_int64 * 6)()
    heavy_seen = 0
    big = ((8, 5, 10, 0, 8.0), (4, 5, 10, 1, 16.0))     # the library's thresholds, 8 ranks, one rank each
    small = ((8, 4, 10, 0, 8.0), (4, 4, 3, 0, 8.0), (3, 3, 8, 1, 20.0))
    combos = [(t, 8, 512, 2048, (0,)) for t in big]
    combos += [(t, nr, light, ps, (0, nr - 1)) for t in small for nr in (2, 3, 8) for light, ps in ((512, 2048), (0, 4), (8, 64))]
    for (M, L, K, nl, thresh), nranks, light, per_share, ranks in combos:
        for rank in ranks:
            rc = emu.hdp_emu_crown_jobs_check(M, L, K, nl, thresh, nranks, rank, light, per_share, out)
            assert rc == 0, (M, L, K, nl, nranks, rank, light, per_share, rc, list(out))
            heavy_seen += out[0]
            if out[0]:
                assert out[1] >= out[0] and out[4] > 0
    assert heavy_seen > 0


def test_crown_partial_pass_by_shares_matches_oracle(H, O, emu):
    """The crown partial pass of a sharded solve by per-node shares (mlp_build_crown_jobs -> prereduce_share ->
    combine_item over the records; on the GPU: mlp_prereduce_kernel with jobs, then mlp_combine_kernel): the share
    sizes are lowered so that the crown nodes of a small tree are split over several shares each, for a local and a
    non-local problem, 2 to 8 emulated ranks, with the error bar.  Same tapes: the oracle's value and wall-hit count;
    the standard error equals the unsharded solve's."""
    emu.hdp_emu_set_crown_jobs.argtypes = [C.c_long, C.c_long]
    emu.hdp_emu_crown_jobs_used.restype = C.c_long
    dt = np.float64
    cases = [
        (H.ParabolicPDEProblem(0.0, 0.7, np.zeros(3), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0), neumann_bc=(-0.5, 0.5)),
         H.MLP(M=4, L=4, K=3), False),
        (H.PIDEProblem(0.0, 0.4, np.zeros(2), (0.0, 0.5), H.ExpSumSq(0.25), H.FNLFisher(), neumann_bc=(-0.5, 0.5)),
         H.MLP(M=3, L=3, K=8, mc_sample=H.UniformSampling(-0.5, 0.5)), True),
    ]
    try:
        for prob, alg, nl in cases:
            tapes = random_tapes(H, prob, alg, dt, seed=11)
            uo, info = oracle_solve(H, O, prob, alg, dt, tapes)
            cn = H.mlp_counts(alg.M, alg.L, alg.K, nonlocal_=nl)
            target = int((cn["path_steps"] + 0.5 * cn["sampler_calls"]) / 8.0)   # a fine cut: many units per crown node
            u1, rc1, st1 = emu_solve(H, emu, prob, alg, dt, tapes=tapes, target_units=target, nranks=1, std_err=True)
            assert rc1 == 0 and rel_err(float(u1[0]), uo) <= 1e-12
            for light, per_share in ((0, 4), (2, 16), (512, 2048)):
                emu.hdp_emu_set_crown_jobs(light, per_share)
                for nranks in (2, 3, 8):
                    before = emu.hdp_emu_crown_jobs_used()
                    u, rc, st = emu_solve(H, emu, prob, alg, dt, tapes=tapes, target_units=target, nranks=nranks, std_err=True)
                    assert rc == 0
                    assert rel_err(float(u[0]), uo) <= 1e-12, (light, per_share, nranks, float(u[0]), uo)
                    assert st.bounces == info["bounces"], (light, per_share, nranks)
                    assert abs(st.std_error - st1.std_error) <= 1e-10 * st1.std_error, (st.std_error, st1.std_error)
                    assert list(st.level_n) == list(st1.level_n)
                    if per_share <= 16:   # the path under test was taken (by every emulated rank)
                        assert emu.hdp_emu_crown_jobs_used() >= before + nranks, (light, per_share, nranks)
    finally:
        emu.hdp_emu_set_crown_jobs(512, 2048)


def alg_work(H, alg):
    cn = H.mlp_counts(alg.M, alg.L, alg.K, nonlocal_=True)
    return cn["path_steps"] + 0.5 * cn["sampler_calls"]
