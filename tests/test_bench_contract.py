"""The JSON line bench.py prints is a contract with the driver: this checks the committed line of the round
(profiles/, produced on a B200 by the command in its `config`) for every key the contract names. CPU only."""
import glob
import json
import os

from cpu_checkers import ROOT


def _latest_line():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r*_final_n1.json")))   # by name: r01 < r01b < r02 ...
    assert files, "no committed bench line under profiles/"
    with open(files[-1]) as f:
        return json.load(f)


def test_committed_bench_line_has_the_contract_keys():
    d = _latest_line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "GCUPS" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert 0 < d["e2e"]["value"] < d["value"]            # host buffers cost something: not a copy of the device number
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_roofline_traffic_comes_from_the_committed_capture():
    d = _latest_line()
    with open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")) as f:
        tr = json.load(f)
    assert d["roofline"]["traffic"] == tr["fill_short_kernel"]["dram_bytes_per_launch"]
    assert abs(d["roofline"]["alu_pipe_frac"] - tr["fill_short_kernel"]["alu_pipe_pct"] / 100.0) < 1e-9
    assert os.path.exists(os.path.join(ROOT, "profiles", "launches_r02_bench.csv"))


def test_round2_fields_of_the_line():
    """What the round-1 review asked the line to carry: the arithmetic that ran, e2e over the driver's step count with the
    box's H2D ceiling beside it, the kernel's own ceiling, and configs 4 / 5 as strong-scaling sections."""
    d = _latest_line()
    assert "int16x2" in d["dtype"]
    assert d["e2e"]["steps"] == d["steps"]
    for k in ("h2d_ceiling_gbs", "h2d_bound_ms_per_step", "frac_of_h2d_bound", "pointer_api"):
        assert k in d["e2e"], k
    assert 0 < d["e2e"]["frac_of_h2d_bound"] <= 1.0
    for k in ("alu_pipe_frac", "issue_slots_per_cell_pair", "ceiling_tcups", "frac_of_kernel_ceiling"):
        assert k in d["roofline"], k
    assert 0 < d["roofline"]["frac_of_kernel_ceiling"] < 1
    c4, c5 = d["extra"]["c4_strong"], d["extra"]["c5_strong"]
    assert c4["reads"] == 100_000 and c4["mapped"] > 99_000 and c4["reads_per_s"] > 0 and "shard.partition" in c4["split"]
    assert c5["pairs"] == 10_000 and c5["gcups"] > 0 and c5["cells"] == 1e12
    assert d["parity_spot_check"]["ok"] is True and "cigar bytes" in d["parity_spot_check"]["fields"]
