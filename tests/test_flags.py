"""Flag-surface parity with the reference."""
import pytest

from distributedmnist_b200.flags import FlagError, FlagValues, define_reference_flags

# name -> default, transcribed from SURVEY §2.2 (reference distributed_train.py:36-99,
# nn_eval.py:36-45, sync_replicas_optimizer_modified.py:38)
REFERENCE_DEFAULTS = {
    "worker_times_cdf_method": False, "interval_method": False, "should_summarize": False,
    "timeline_logging": False, "job_name": "", "ps_hosts": "", "worker_hosts": "",
    "train_dir": "/tmp/imagenet_train", "rpc_port": 1235, "save_results_period": 1000,
    "max_steps": 1000000, "drop_connect": False, "batch_size": 128, "subset": "train",
    "log_device_placement": False, "task_id": 0, "num_replicas_to_aggregate": -1,
    "save_interval_secs": 20, "save_summaries_secs": 300, "initial_learning_rate": 0.1,
    "num_epochs_per_decay": 2.0, "learning_rate_decay_factor": 0.999, "drop_connect_probability": 0.9,
    "interval_ms": 1000, "eval_dir": "/tmp/imagenet_eval", "checkpoint_dir": "/tmp/imagenet_train",
    "eval_interval_secs": 1, "run_once": False,
}


def test_every_reference_flag_exists_with_same_default():
    F = define_reference_flags(FlagValues())
    d = F.defaults_dict()
    for name, default in REFERENCE_DEFAULTS.items():
        assert name in d, name
        assert d[name] == default and type(d[name]) is type(default), name


def test_parse_styles():
    F = define_reference_flags(FlagValues())
    rest = F.parse(["--batch_size=64", "--initial_learning_rate", ".0008", "--interval_method=true",
                    "--worker_times_cdf_method=False", "--drop_connect", "--norun_once",
                    "--worker_hosts='a:1234,b:1234'", "positional"])
    assert rest == ["positional"]
    assert F.batch_size == 64 and F.initial_learning_rate == pytest.approx(0.0008)
    assert F.interval_method is True and F.worker_times_cdf_method is False
    assert F.drop_connect is True and F.run_once is False
    assert F.worker_hosts == "a:1234,b:1234"   # cfg-style quotes stripped
    assert F.is_present("batch_size") and not F.is_present("max_steps")


def test_reference_train_command_line_parses(fresh_flags):
    # The command template of the reference's cfg files (cfg/50_workers/*:74-87).
    argv = ("--batch_size=128 --initial_learning_rate=.0008 --learning_rate_decay_factor=1 "
            "--num_epochs_per_decay=1.0 --train_dir=/tmp/x/train_dir --worker_hosts=h0:1234,h1:1234 "
            "--ps_hosts=p:1234 --task_id=1 --num_replicas_to_aggregate=10 --job_name=worker "
            "--interval_method=true --interval_ms=3000 --timeline_logging=false").split()
    assert fresh_flags.parse(argv) == []
    assert fresh_flags.interval_ms == 3000 and fresh_flags.task_id == 1 and fresh_flags.job_name == "worker"


def test_errors():
    F = define_reference_flags(FlagValues())
    with pytest.raises(FlagError):
        F.parse(["--no_such_flag=1"])
    with pytest.raises(FlagError):
        F.parse(["--batch_size=abc"])
    with pytest.raises(FlagError):
        F.parse(["--run_once=maybe"])
    assert F.parse(["--mystery=1"], known_only=True) == ["--mystery=1"]
