"""Host-side helpers of bench.py (no GPU): the instruction-issue roofline annotation."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_issue_roofline_uses_the_committed_capture(tmp_path):
    b = _bench()
    tab = tmp_path / "instructions.json"
    tab.write_text(json.dumps({"workload": "C2", "pack_classify_sketch": 1_000_000_000}))
    r = b.issue_roofline("pack_classify_sketch", "C2", 2e-3, 148, 2000.0, path=str(tab))
    assert r["peak"] == 148 * 4 * 2000.0e6
    assert abs(r["achieved"] - 5e11) < 1 and abs(r["frac"] - 5e11 / r["peak"]) < 1e-4
    # another workload, an unknown kernel or a missing launch time give no figure instead of a wrong one
    assert b.issue_roofline("pack_classify_sketch", "C1", 2e-3, 148, 2000.0, path=str(tab)) is None
    assert b.issue_roofline("consensus", "C2", 2e-3, 148, 2000.0, path=str(tab)) is None
    assert b.issue_roofline("pack_classify_sketch", "C2", 0.0, 148, 2000.0, path=str(tab)) is None


def test_committed_instruction_table_is_for_the_bench_default():
    with open(os.path.join(ROOT, "profiles", "instructions.json")) as f:
        tab = json.load(f)
    assert tab["workload"] == "C2" and tab["classify_sketch_packed"] > 1e9 and tab["consensus"] > 5e8
