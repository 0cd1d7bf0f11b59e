"""The per-kernel table bench.py reports as `roofline` comes from profiles/r02_issue_<config>.json, written by
profiles/summarize_ncu.py from an ncu capture; these tests pin the parts that are easy to get silently wrong."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _summarize():
    spec = importlib.util.spec_from_file_location("summarize_ncu", os.path.join(ROOT, "profiles", "summarize_ncu.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_state_rows_from_launch_shared_memory():
    s = _summarize()
    chain = "void mbsv::mbsv_chain_kernel<{}>(mbsv::DevProblem, mbsv::DevRun)"
    # 32-bit tiles: 4 warps x rows x 128 bytes; SMTOT instantiations keep 2560 + 4096 * MAXE bytes behind the tiles
    assert s.state_rows(chain.format("1, 0, 1, 2, 0, 0, 3, 0, int"), 4 * 128 * 128) == 128
    assert s.state_rows(chain.format("1, 0, 1, 2, 0, 0, 0, 1, int"), 4 * 48 * 128 + 2560 + 8192) == 48
    # 8-bit tiles (kern_fast8.cu): 4 warps x rows x 32 bytes
    assert s.state_rows(chain.format("1, 0, 1, 2, 0, 0, 0, 1, signed char"), 4 * 96 * 32 + 2560 + 8192) == 96
    # the global-workspace class has no tile, only the counters and totals
    assert s.state_rows(chain.format("1, 0, 0, 2, 0, 0, 0, 1, int"), 2560 + 8192) == 0
    assert s.state_rows("void mbsv::mbsv_memo_kernel<0, 0>(mbsv::DevProblem, mbsv::DevRun)", 8192) == 16
    assert s.state_rows("void mbsv::mbsv_giant_kernel<512, 2, 2>(mbsv::DevProblem, mbsv::DevRun)", 122688) == 0


def test_committed_issue_tables_have_what_bench_reads():
    for cfg in ("cfg4", "cfg5"):
        d = json.load(open(os.path.join(ROOT, "profiles", f"r02_issue_{cfg}.json")))
        dom = d["dominant"]
        assert {"kernel", "issue_active_pct", "active_lanes", "dram_bytes_per_launch"} <= set(dom)
        assert 0 < dom["issue_active_pct"] <= 100 and 0 < dom["active_lanes"] <= 32
        assert abs(dom["frac"] - dom["issue_active_pct"] * dom["active_lanes"] / 3200.0) < 1e-9
        shares = [k["share_of_kernel_time"] for k in d["kernels"]]
        assert abs(sum(shares) - 1.0) < 1e-6 and dom["share_of_kernel_time"] == max(shares)
