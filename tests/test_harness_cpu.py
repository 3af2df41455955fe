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
    assert "torch.distributed.run" not in one and one[-1] == "--no-cpu-baseline" and "--only-headline" in one
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


spec2 = importlib.util.spec_from_file_location("partitioning", os.path.join(ROOT, "benchmarks", "partitioning.py"))
pt = importlib.util.module_from_spec(spec2)
spec2.loader.exec_module(pt)


def test_partitioning_harness_layout_and_scraping(capsys):
    """benchmarks/partitioning.py: the reference's column layout (benchmark-partitioning.sh:81-95), its flag order, and
    the scraping of the SECOND `Elapsed wall clock time` line."""
    cols = pt.header(2).split()
    assert cols[:5] == ["#PARTITIONS", "INS_PPPCSR0", "INS_PPPCSR1", "INS_PPPCSR_Avg", "INS_PPPCSR_Stddev"]
    assert cols[5:9] == ["DEL_PPPCSR0", "DEL_PPPCSR1", "DEL_PPPCSR_Avg", "DEL_PPPCSR_Stddev"] and len(cols) == 17
    cmd = pt.cli_command("exe", "-pppcsrnuma", True, 8, 1000, "c.bin", "u.bin", 4, [])
    assert cmd.index("-delete") < cmd.index("-update_file=u.bin") and cmd.index("-size=1000") < cmd.index("-update_file=u.bin")
    assert cmd[-1] == "-partitions_per_domain=4"
    out = "Elapsed wall clock time: 12\nfoo\nElapsed wall clock time: 3\n" + json.dumps({"updates": 5, "device_ms": 0.25}) + "\n"
    assert pt.scrape(out) == (3.0, 0.25)
    assert pt.main(["--partitions", "1", "2", "--reps", "2", "--dry-run"]) == 0
    lines = capsys.readouterr().out.strip().splitlines()
    assert lines[0].startswith("#PARTITIONS") and [l.split()[0] for l in lines[1:]] == ["1", "2"]
    assert all(len(l.split()) == len(lines[0].split()) for l in lines[1:])
