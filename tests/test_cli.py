"""mcmc-multispde_b200/twoDsim.py's option schema against the REFERENCE'S OWN parser (CPU).

tests/golden/ref_cli_args.json holds what /root/reference/twoDsim.py:68-99 (+ parser_help.py) parsed from a list of command
lines (oracle/make_cli_golden.py).  The product's parser must give every option the reference's driver reads the same value,
reject what the reference rejects, and -- its documented extensions -- accept --basis-n, --result-dir and the --no<flag> form
for hyphenated flags (which the reference parses into an attribute nothing reads)."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

with open(os.path.join(ROOT, "tests", "golden", "ref_cli_args.json")) as _fh:
    GOLD = json.load(_fh)


@pytest.fixture(scope="module")
def cli():
    spec = importlib.util.spec_from_file_location("twoDsim_b200", os.path.join(ROOT, "mcmc-multispde_b200", "twoDsim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("entry", GOLD["accepted"], ids=lambda e: " ".join(e["argv"]) or "defaults")
def test_options_parse_like_the_reference(cli, entry):
    ours = vars(cli.parse_args(entry["argv"]))
    ref = {k: v for k, v in entry["args"].items() if "-" not in k}      # hyphenated keys: the reference's unread --no<flag> dests
    assert set(ref) <= set(ours)
    for k, v in ref.items():
        assert ours[k] == v and type(ours[k]) is type(v), (k, ours[k], v)
    assert set(ours) - set(ref) == {"result_dir", "distributed"}         # the options added here


@pytest.mark.parametrize("argv", GOLD["rejected"], ids=lambda a: " ".join(a))
def test_rejects_what_the_reference_rejects(cli, argv, capsys):
    with pytest.raises(SystemExit):
        cli.parse_args(argv)
    capsys.readouterr()


def test_extensions(cli):
    a = cli.parse_args(["--basis-n", "96", "--result-dir", "/tmp/out", "--noenable-step-adaptation", "--noprint-progress"])
    assert a.n == 96 and a.result_dir == "/tmp/out"
    assert a.enable_step_adaptation is False and a.print_progress is False and a.distributed is False
    assert cli.parse_args(["--distributed", "--basis-n", "96"]).distributed is True
