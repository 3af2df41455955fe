"""Host-side logic of the strong-scaling harness (benchmarks/strong_scaling.py): command lines, scraping of
bench.py's JSON line and the CSV layout of the reference's benchmark-strong-scaling.sh (header columns, mean and
sample standard deviation)."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("strong_scaling", os.path.join(ROOT, "benchmarks", "strong_scaling.py"))
ss = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ss)


def test_commands_follow_the_driver_launch_contract():
    one = ss.bench_command(1, "insert", 24, 100, 3, 3, 29541)
    assert "torch.distributed.run" not in one and one[-1] == "--no-cpu-baseline" and "--strong" in one
    four = ss.bench_command(4, "delete", 24, 100, 3, 3, 29541)
    assert four[1:3] == ["-m", "torch.distributed.run"] and "--nproc-per-node" in four
    assert four[four.index("--master-addr") + 1] == "127.0.0.1"
    assert four[four.index("--gpus") + 1] == "4" and four[four.index("--workload") + 1] == "delete"


def test_scrape_and_statistics():
    out = "NCCL banner\n" + json.dumps({"metric": "edge_updates_per_sec", "ms_per_step": 1.25}) + "\n"
    assert ss.scrape_ms(out) == 1.25
    a, sd = ss.avg_stddev([1.0, 2.0, 3.0])
    assert a == 2.0 and abs(sd - 1.0) < 1e-12  # sample standard deviation, as the reference's awk line
    assert ss.avg_stddev([5.0]) == (5.0, 0.0)
    cols = ss.header(2).split()
    assert cols == ["#GPUS", "INS_SHARDS0", "INS_SHARDS1", "INS_SHARDS_Avg", "INS_SHARDS_Stddev",
                    "DEL_SHARDS0", "DEL_SHARDS1", "DEL_SHARDS_Avg", "DEL_SHARDS_Stddev"]


def test_dry_run_prints_one_row_per_gpu_count(capsys):
    assert ss.main(["--gpus", "1", "2", "--reps", "2", "--dry-run"]) == 0
    lines = capsys.readouterr().out.strip().splitlines()
    assert lines[0].startswith("#GPUS") and [l.split()[0] for l in lines[1:]] == ["1", "2"]
    assert all(len(l.split()) == len(lines[0].split()) for l in lines[1:])
