"""Distributed plumbing on CPU/gloo, world_size=2 (BASELINE.json config 1):
2-layer MLP sync-SGD, replicas bit-identical after each step, K-of-N masking, the three modes."""
import json
import os
import re

import pytest

from distributedmnist_b200.parallel.launcher import run_replicas

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "_gloo_worker.py")
COMMON = ["--model=mlp2", "--mlp_hidden=32", "--batch_size=16", "--backend=gloo", "--initial_learning_rate=0.05",
          "--save_interval_secs=1000"]


def _run(tmp_path, extra, n=2, timeout=180):
    out = str(tmp_path / "out")
    argv_common = COMMON + ["--train_dir=%s" % (tmp_path / "train_dir")] + extra
    # each rank writes its own json: pass a template resolved through RANK in the worker's argv
    procs_argv = [WORKER, str(tmp_path / "res_RANK.json")] + argv_common
    # the launcher passes identical argv to every rank; the worker substitutes RANK itself
    codes = run_replicas(["-c", "import os,sys,runpy; sys.argv=[a.replace('RANK', os.environ['RANK']) for a in sys.argv[1:]];"
                          " runpy.run_path(sys.argv[0], run_name='__main__')"] + procs_argv,
                         n, timeout=timeout, out_dir=out)
    logs = [open(os.path.join(out, f)).read() for f in sorted(os.listdir(out))]
    assert codes == [0] * n, "\n".join(l[-2000:] for l in logs)
    res = [json.load(open(tmp_path / ("res_%d.json" % r))) for r in range(n)]
    return res, logs


def test_sync_sgd_replicas_bit_identical_and_loss_decreases(tmp_path):
    res, logs = _run(tmp_path, ["--max_steps=30"])
    a, b = res
    assert a["fingerprints"] == b["fingerprints"] and len(a["fingerprints"]) == 31
    assert len(set(a["fingerprints"])) == 31                       # parameters actually move
    assert all(i[1] and i[2] == 0b11 and i[3] == 2 for i in a["infos"])
    assert a["final_fp"] == b["final_fp"] and a["final_step"] == 31
    assert sum(a["losses"][-5:]) < sum(a["losses"][:5])
    # reference log-line format, as scraped by tools/benchmark.py (".*step ([0-9]*),.*")
    master = [l for l in logs if "Worker 0:" in l][0]
    steps = [int(m) for m in re.findall(r"Worker 0: .*: step ([0-9]*), loss = [0-9.]+, train_acc = [0-9.]+, "
                                        r"test_acc = 0.000000\([0-9.]+ examples/sec; [0-9.]+  sec/batch\)", master)]
    assert steps == list(range(1, 31))
    assert "Elapsed Time:" in master and "Global step attained: 31" in master
    # final checkpoint with the reference's naming
    assert os.path.exists(tmp_path / "train_dir" / "checkpoint")
    assert os.path.exists(tmp_path / "train_dir" / "model.ckpt-31.index")


def test_k_of_n_masks_the_straggler(tmp_path):
    # rank 1 always arrives 30 ms late; K=1 -> rank 0's gradient alone commits every step
    res, _ = _run(tmp_path, ["--max_steps=12", "--num_replicas_to_aggregate=1", "--inject_straggler=1:1.0:30000"])
    a, b = res
    assert a["fingerprints"] == b["fingerprints"]                   # still identical replicas
    assert all(i[2] == 0b01 and i[3] == 1 for i in a["infos"])      # same mask on both, divisor = 1
    assert all(i[1] for i in a["infos"]) and not any(i[1] for i in b["infos"])
    assert a["accepted"] == 13 and b["dropped"] == 13


def test_cdf_mode_full_barrier_with_timing_lines(tmp_path):
    res, logs = _run(tmp_path, ["--max_steps=60", "--worker_times_cdf_method=true", "--interval_method=false"])
    a, b = res
    assert a["fingerprints"] == b["fingerprints"]
    master = [l for l in logs if "Worker 0:" in l][0]
    m = re.findall(r"ELAPSED TIMES (.*)", master)
    assert m, "no ELAPSED TIMES line"
    entries = eval(m[-1])                        # tools/benchmark.py parses it exactly like this
    assert all(len(e) == 3 and e[2] > 10 for e in entries)
    assert {e[1] for e in entries} == {0, 1}     # both workers reported, tables not aliased
    assert re.findall(r"ITERATION TIMES \[", master)


def test_interval_mode_applies_mean_of_whatever_arrived(tmp_path):
    res, _ = _run(tmp_path, ["--max_steps=5", "--interval_method=true", "--interval_ms=150"])
    a, b = res
    assert a["final_fp"] == b["final_fp"] and a["final_step"] == 6


def test_interval_and_cdf_flags_together_do_not_deadlock(tmp_path):
    """Both mode flags set: interval wins (reference distributed_train.py:179-183) and the cdf telemetry -- whose table
    exchange is a collective keyed on the LOCAL iteration -- must stay off, because interval replicas are not in lock step."""
    res, logs = _run(tmp_path, ["--max_steps=6", "--interval_method=true", "--interval_ms=100",
                                "--worker_times_cdf_method=true"], timeout=120)
    a, b = res
    assert a["final_fp"] == b["final_fp"] and a["final_step"] >= 7
    assert not any("ELAPSED TIMES" in l for l in logs)


def test_fd_channel_passes_descriptors_between_ranks(tmp_path):
    """The fd exchange behind the VMM / NVLS-multicast allocator (parallel/symm_mem.py): SCM_RIGHTS over abstract AF_UNIX
    datagram sockets, all-to-all and broadcast, plus the collective AND that keeps the ranks on the same branch."""
    worker = os.path.join(HERE, "_fdchannel_worker.py")
    n = 3
    codes = run_replicas([worker, str(tmp_path / "res_RANK.json")], n, timeout=120, out_dir=str(tmp_path / "out"))
    logs = "\n".join(open(os.path.join(tmp_path, "out", f)).read()[-1500:] for f in sorted(os.listdir(tmp_path / "out")))
    assert codes == [0] * n, logs
    res = [json.load(open(tmp_path / ("res_%d.json" % r))) for r in range(n)]
    for r in range(n):
        for row in res[r]["texts"]:
            assert row == ["self" if q == r else "hello from rank %d" % q for q in range(n)], row
        assert res[r]["bcast"] == "hello from rank 0"
        assert res[r]["all_ok"] is False
