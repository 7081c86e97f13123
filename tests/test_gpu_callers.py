"""GPU: the vision call sequence of the reference's callers on the drop-in modules with the real engine.
  * everywhere: the recorded sequence of GameSession.on_frame / calibrate_sensitivity's loop (tests/golden/callers.json,
    made by tools/make_golden_callers.py from the UNMODIFIED callers on the reference's own modules) is replayed and
    must give the same occupancy sets, visual changes and ChangeDetector dictionaries;
  * where the reference checkout exists: the unchanged callers themselves (game_session, play_lichess,
    calibrate_sensitivity) run on the real engine and must agree with the reference run."""
import importlib
import json
import os
import sys

import pytest

import chessboard_vision_b200.dropin as dropin

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import caller_harness as ch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture()
def mods(engine, monkeypatch):
    pytest.importorskip("cv2")
    import chessboard_vision_b200.engine as engine_mod
    monkeypatch.setitem(engine_mod._default, 0, engine)
    monkeypatch.syspath_prepend(dropin.PATH)
    out = {}
    for name in ch.VISION:
        sys.modules.pop(name, None)
        out[name] = importlib.import_module(name)
    yield out
    for name in ch.VISION:
        sys.modules.pop(name, None)


def test_game_session_sequence_replayed(mods):
    golden = json.load(open(os.path.join(G, "callers.json")))["game_session"]
    g = ch.replay_game(mods, ch.scenario_game(), golden)
    assert [[list(p) for p in o] for o in g["occupied"]] == golden["occupied"]
    assert [[list(p) for p in o] for o in g["changes"]] == golden["changes"]


def test_calibrate_sensitivity_sequence_replayed(mods):
    golden = json.load(open(os.path.join(G, "callers.json")))["calibrate_sensitivity"]
    s = ch.replay_sensitivity(mods, ch.scenario_sensitivity(), golden["loop_turns"])
    norm = lambda d: [{k: list(v) for k, v in x.items()} for x in d]
    assert norm(s["detailed"]) == norm(golden["detailed"]) and s["changes"] == golden["changes"]
    assert [list(p[:2]) + [[list(q) for q in p[2]]] for p in s["patterns"]] == \
           [list(p[:2]) + [[list(q) for q in p[2]]] for p in golden["patterns"]]


@pytest.mark.skipif(not ch.have_reference(), reason="needs the reference checkout at /root/reference")
def test_unchanged_callers_on_the_real_engine(engine):
    frames = ch.scenario_game()
    ref, got = ch.run_game_session("reference", frames), ch.run_game_session("dropin", frames, engine)
    assert got["vision"] == "dropin" and got["occupied"] == ref["occupied"] and got["moves"] == ref["moves"] == ["e2e4"]
    frames = ch.scenario_sensitivity()
    ref, got = ch.run_calibrate_sensitivity("reference", frames), ch.run_calibrate_sensitivity("dropin", frames, engine)
    assert got["detailed"] == ref["detailed"] and got["patterns"] == ref["patterns"]
