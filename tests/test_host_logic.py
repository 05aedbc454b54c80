"""Host-side logic that needs no GPU: the state ranges of the pipelined host-buffer iteration."""
import numpy as np
import pytest

from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel
from variational_gradient_matching_for_dynamical_systems_b200.engine import LocalProblem, pipeline_ranges


def _problem(K, hidden):
    return LocalProblem.from_model(OdeModel.lorenz96(K), hidden)


@pytest.mark.parametrize("K,n", [(96, 2), (96, 4), (1000, 3), (100000, 4), (4096, 8)])
def test_pipeline_ranges_cover_everything_and_keep_the_snapshot(K, n):
    hidden = np.arange(1, K, 2)
    prob = _problem(K, hidden)
    pr = pipeline_ranges(prob, n)
    assert pr is not None and pr["n"] == n
    kp, sp, op, rp, hx, hd = pr["kp"], pr["sp"], pr["op"], pr["rp"], pr["hx"], pr["hd"]
    hp = prob.host_plan
    # Lorenz96: ODE k reads states k-2..k+1, a state occurs in ODEs u-1..u+2
    assert (hd, hx) == (2, 4)
    assert kp[0] == 0 and kp[-1] == K and np.all(kp % 8 == 0) and np.all(np.diff(kp) >= 2 * hx + 8)
    # staged slot ranges partition the hidden slots; solve ranges + seam partition dependency level 0
    assert sp[0] == 0 and sp[-1] == hp.n_hidden and np.all(np.diff(sp) > 0)
    n0 = hp.level_ptr[1]
    assert rp[-1] == n0 and np.all(np.diff(rp) >= 0)
    lvl0 = np.asarray(hp.level_slots[:n0])
    solved = np.concatenate([lvl0[rp[c]:rp[c + 1]] for c in range(n)] + [lvl0[:rp[0]]])
    assert np.array_equal(np.sort(solved), lvl0)
    # every later-level slot is copied out with the last range
    later = np.asarray(hp.level_slots[n0:])
    assert len(later) == 0 or later.min() >= op[-2]
    # snapshot safety: when range c is staged (c >= 1), the systems solved so far are those of the solve ranges < c;
    # none of their states may lie within the reach of range c's coefficients (cyclically)
    state_of = np.asarray(hp.slot_state)
    for c in range(1, n):
        done_states = state_of[np.concatenate([lvl0[rp[q]:rp[q + 1]] for q in range(c)])]
        lo, hi = kp[c] - hx, kp[c + 1] + hx            # states range c's staging may read
        reach = np.arange(lo, hi) % K
        assert not np.intersect1d(done_states, reach).size, c
    assert np.array_equal(pr["last_rows"], np.concatenate([lvl0[rp[-2]:], lvl0[:rp[0]]]))
    # the seam (solved last) is exactly the hidden states below hx
    assert np.array_equal(state_of[lvl0[:rp[0]]], hidden[hidden < hx])
    # a solve range only needs coefficients that were staged by then
    for c in range(n):
        st = state_of[lvl0[rp[c]:rp[c + 1]]]
        assert st.size == 0 or st.max() < kp[c + 1]


def test_pipeline_ranges_with_uneven_fractions():
    K = 100000
    prob = _problem(K, np.arange(1, K, 2))
    pr = pipeline_ranges(prob, (0.15, 0.55, 0.9, 1.0))
    assert pr["n"] == 4 and list(pr["kp"]) == [0, 15000, 55000, 90000, 100000]
    assert pr["rp"][-1] == prob.host_plan.level_ptr[1] and np.all(np.diff(pr["rp"]) > 0)
    with pytest.raises(ValueError):
        pipeline_ranges(prob, (0.5, 0.4, 1.0))
    with pytest.raises(ValueError):
        pipeline_ranges(prob, (0.5, 0.9))


def test_pipeline_ranges_refuse_what_they_cannot_handle():
    K = 96
    prob = _problem(K, np.arange(1, K, 2))
    assert pipeline_ranges(prob, 1) is None
    assert pipeline_ranges(prob, 4, theta_mode="proxy") is None          # theta is needed before the sweep
    assert pipeline_ranges(prob, 16) is None                             # ranges shorter than two halos
    assert pipeline_ranges(_problem(K, np.arange(K)), 4) is None         # all-hidden chain: every level depends on the last
    assert pipeline_ranges(_problem(K, np.array([5, 3, 1])), 2) is None  # slots not in ascending state order
