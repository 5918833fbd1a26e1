"""The roofline arithmetic of bench.py (algorithmic bytes / flops per launch) against SURVEY.md 8(d)'s figures -- no GPU."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_mlp_flops_match_survey_8d():
    costs = _bench().stage_costs(8, "bf16")
    sa = [costs["layer%d:mlp" % i][1] / 1e9 for i in range(1, 5)]
    fp = [costs["fa_layer%d:mlp" % i][1] / 1e9 for i in range(1, 5)]
    for got, exp in zip(sa, (3.42, 4.35, 4.32, 4.31)):   # grouped MLP + max
        assert abs(got - exp) < 0.01, (got, exp)
    for got, exp in zip(fp, (0.54, 1.34, 3.76, 25.97)):  # 0.5 + 1.3 + 3.8 + 26.0 = 31.6 GFLOP
        assert abs(got - exp) < 0.01, (got, exp)
    assert all(costs[k][0] == "tensor" for k in costs if k.endswith(":mlp"))


def test_search_and_gather_bytes():
    b = _bench()
    bf, f32 = b.stage_costs(8, "bf16"), b.stage_costs(8, "fp32")
    n, m, k = 32768, 2048, 32
    # FPS: B*(12n + 4m); SA1 in the bf16 path gathers inside the chain, so its search stage moves the cloud, the queries and
    # writes idx / pts_cnt only
    assert bf["layer1:fps"][:2] == ("hbm", 8 * (12 * n + 4 * m))
    assert bf["layer1:ballquery_group"][:2] == ("hbm", 8 * (12 * n + 12 * m + 4 * m * k + 4 * m))
    # pair evaluations of the reference's scans: FPS B*(m-1)*n, ball query / three_nn B*n*m
    assert bf["layer1:fps"][2] == 8 * (m - 1) * n and bf["layer1:ballquery_group"][2] == 8 * m * n and bf["fa_layer4:three_nn"][2] == 8 * n * m
    # fp32 path: the grouped (m, K, 3 + C) rows are written as fp32 (SURVEY.md 8(d) formula with e_out = 4, ld = c + 3)
    assert f32["layer1:ballquery_group"][1] == 8 * (12 * n + 12 * m + n * 3 * 4 + 4 * m * k + 4 * m + m * k * 6 * 4)
    # SA2 (C = 64): tile image of ld = 128 bf16 columns
    assert bf["layer2:ballquery_group"][1] == 8 * (12 * 2048 + 12 * 512 + 2048 * 64 * 4 + 4 * 512 * 32 + 4 * 512 + 512 * 32 * 128 * 2)
    # three_nn FP4: B*(12n + 12m + 36n)
    assert bf["fa_layer4:three_nn"][:2] == ("hbm", 8 * (12 * n + 12 * m + 36 * n))
    # the split image of the default precision stores a [hi | lo] pair per element: 4 bytes
    x3 = b.stage_costs(8, "bf16x3")
    assert x3["layer2:ballquery_group"][1] - bf["layer2:ballquery_group"][1] == 8 * 512 * 32 * 128 * 2
    assert x3["layer1:mlp"] == bf["layer1:mlp"]
    # issued MMA flops per bf16 pass: 16-column k-slices (SA1: 6 -> 16 input columns); the commuted fa_layer4 issues layers 1.. only
    assert bf["layer1:mlp"][3] == 2 * 8 * m * k * (16 * 32 + 32 * 32 + 32 * 64)
    assert bf["fa_layer4:mlp"][3] == 2 * 8 * n * (128 * 128 + 128 * 128) and bf["fa_layer4:interpolate"][0] == "tensor"
    assert b.stage_costs(8, "fp32")["fa_layer4:interpolate"][0] == "hbm"


def test_peaks_loader_prefers_measured_file(tmp_path, monkeypatch):
    b = _bench()
    peaks = b.load_peaks()
    assert peaks["hbm_gbs"] > 1000 and peaks["bf16_tflops_sustained"] <= peaks["bf16_tflops"] * 1.01
    assert peaks["source"] in ("measured", "fallback")


def test_bench_forwards_unknown_flags_unabbreviated():
    """Flags bench.py does not know go to tools/bench_<workload>.py; a prefix of one of bench.py's own flags must not be eaten on the
    way (argparse abbreviations: --no-graph would have matched --no-graphs)."""
    import bench
    ap = bench.make_parser()
    args, rest = ap.parse_known_args(["--workload", "cfg4", "--steps", "3", "--nccl-moments", "--eager-launch", "--no-graph"])
    assert args.workload == "cfg4" and args.steps == 3 and not args.no_graphs
    assert rest == ["--nccl-moments", "--eager-launch", "--no-graph"]
    args, rest = ap.parse_known_args(["--no-graphs", "--depth", "4"])
    assert args.no_graphs and args.depth == 4 and rest == []
