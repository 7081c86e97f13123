"""Record what the reference's UNCHANGED callers see on the synthetic camera scenarios of tests/caller_harness.py,
running on the reference's OWN vision modules: tests/golden/callers.json.  Build container only (/root/reference).

    python tools/make_golden_callers.py

The GPU box has no reference checkout: there tests/test_gpu_callers.py replays the same vision call sequence on the
drop-in modules and compares with this file; where the checkout exists the callers themselves run on the drop-ins."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import caller_harness as ch


def main():
    game = ch.run_game_session("reference", ch.scenario_game())
    sens = ch.run_calibrate_sensitivity("reference", ch.scenario_sensitivity(), loop_turns=40)
    assert game["vision"] == sens["vision"] == "reference"
    out = {"game_session": {k: game[k] for k in ("occupied", "changes", "noise", "moves", "to_check", "updates", "frames")},
           "calibrate_sensitivity": {"detailed": sens["detailed"], "changes": sens["changes"], "patterns": sens["patterns"],
                                     "loop_turns": 40}}
    path = os.path.join(ROOT, "tests", "golden", "callers.json")
    json.dump(out, open(path, "w"), default=float)
    print("wrote", path, os.path.getsize(path), "bytes; moves", game["moves"], "frames", game["frames"])


if __name__ == "__main__":
    main()
