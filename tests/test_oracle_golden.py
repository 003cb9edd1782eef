"""CPU: the oracle restatement (oracle/mpti_oracle.py) against the golden vectors that were
produced by the reference's own modules (oracle/make_golden.py)."""
import torch

from oracle import mpti_oracle as O
from r3dfsseg_b200.episodes import make_episode
from tests.helpers import knn_sets_match


def test_knn_matches_reference(golden_dgcnn):
    g = golden_dgcnn
    for key in ("x9", "x64"):
        x = g[key]
        idx = O.knn(x, 20)
        ok, bad = knn_sets_match(idx, g["knn_" + key], O.knn_scores(x), largest=True)
        assert ok, f"{key}: {bad} rows differ beyond ties"


def test_edge_feature_matches_reference(golden_dgcnn):
    g = golden_dgcnn
    e = O.get_edge_feature(g["x9"], 20, g["knn_x9"])
    assert torch.equal(e, g["edge_x9"])


def test_dgcnn_matches_reference(golden_dgcnn, fixture_sd):
    g = golden_dgcnn
    l1, l2 = O.dgcnn_forward(g["dgcnn_x"], fixture_sd)
    assert torch.allclose(l1, g["dgcnn_l1"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(l2, g["dgcnn_l2"], rtol=1e-5, atol=1e-6)


def test_fps_count_rule():
    # ceil(fp32(n) * fp32(k/n)) is k or k+1
    ks = {O.fps_count(n, 100) for n in range(101, 4000)}
    assert ks == {100, 101}
    assert all(O.fps_count(n, 4) == 4 for n in range(5, 3000))


import pytest


@pytest.mark.parametrize("name", ["s3dis_2way_1shot", "s3dis_2way_5shot_mdns",
                                  "s3dis_2way_5shot_noisy_mdns", "scannet_3way_5shot_ood_mdns"])
def test_episode_matches_reference(golden_episodes, fixture_sd, name):
    c = golden_episodes[name]
    ep = make_episode(c["seed"], c["n_way"], c["k_shot"], dataset=c["dataset"],
                      noise_ratio=c["noise_ratio"])
    with torch.no_grad():
        out = O.forward_episode(fixture_sd, ep.support_x, ep.support_y, ep.query_x, ep.query_y,
                                eval_mdns=c["eval"])
    ref = c["query_pred"]
    assert out["num_prototypes"] == c["num_prototypes"]
    if c["eval"]:
        assert torch.equal(out["clean_flag"], c["clean_flag"])
    err = (out["query_pred"] - ref).abs().max() / ref.abs().max()
    assert err < 1e-3, err
    agree = (out["query_pred"].argmax(1) == ref.argmax(1)).float().mean()
    assert agree >= 0.999, agree
    assert abs(float(out["loss"]) - float(c["loss"])) < 1e-4
