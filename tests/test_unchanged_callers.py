"""The reference's UNCHANGED callers on the drop-in modules (BASELINE.json north_star: "game_session and
play_lichess run unchanged").  game_session.GameSession (on_calibration_requested + on_frame), play_lichess.main and
calibrate_sensitivity.main are imported from the reference checkout and driven by a fake camera; the run on the
drop-ins (kernels replaced by the oracle-backed fake engine: there is no GPU here) must see exactly what the same run
on the reference's own vision modules sees -- occupancy per frame, visual changes, noise states, the inferred move,
the ChangeDetector dictionaries.  Skipped where the checkout is absent (the GPU box): tests/test_gpu_callers.py
replays the recorded call sequence there."""
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import caller_harness as ch

pytestmark = pytest.mark.skipif(not ch.have_reference(), reason="needs the reference checkout at /root/reference")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture()
def fake(oracle):
    pytest.importorskip("cv2")
    from fake_engine import FakeEngine
    return FakeEngine()


def test_game_session_on_frame(fake):
    frames = ch.scenario_game()
    ref = ch.run_game_session("reference", frames)
    got = ch.run_game_session("dropin", frames, fake)
    assert ref["vision"] == "reference" and got["vision"] == "dropin"          # the callers really imported our modules
    assert got["frames"] == ref["frames"] >= 10
    assert got["occupied"] == ref["occupied"] and got["changes"] == ref["changes"] and got["noise"] == ref["noise"]
    assert got["to_check"] == ref["to_check"] and got["updates"] == ref["updates"]
    assert got["moves"] == ref["moves"] == ["e2e4"]                            # the move is inferred from our occupancy sets
    assert any(c for c in ref["changes"]) and len({len(o) for o in ref["occupied"]}) >= 1
    golden = json.load(open(os.path.join(G, "callers.json")))["game_session"]
    assert [[list(p) for p in o] for o in ref["occupied"]] == golden["occupied"] and ref["moves"] == golden["moves"]


def test_play_lichess_main(fake):
    frames = ch.camera_frames([ch.START] * 30 + [ch.AFTER_E2E4] * 60, seed=2)
    ref = ch.run_play_lichess("reference", frames, loop_turns=70)
    got = ch.run_play_lichess("dropin", frames, fake, loop_turns=70)
    assert got["vision"] == "dropin" and got["on_frame_calls"] == ref["on_frame_calls"] >= 30
    assert got["sent"] == ref["sent"] == ["e2e4"]                              # sent to the (mock) Lichess client
    assert got["shown"] == ref["shown"]


def test_calibrate_sensitivity_main(fake):
    frames = ch.scenario_sensitivity()
    ref = ch.run_calibrate_sensitivity("reference", frames, loop_turns=40)
    got = ch.run_calibrate_sensitivity("dropin", frames, fake, loop_turns=40)
    assert got["vision"] == "dropin" and len(got["detailed"]) == len(ref["detailed"]) >= 8
    assert got["detailed"] == ref["detailed"] and got["changes"] == ref["changes"] and got["patterns"] == ref["patterns"]
    assert any(d for d in ref["detailed"]) and any(p[0] for p in ref["patterns"])      # a move and a "hand" were seen


def test_replay_matches_the_callers(fake, monkeypatch):
    """The restated call sequence the GPU box uses (no checkout there) gives what the callers themselves gave."""
    import importlib
    import chessboard_vision_b200.engine as engine_mod
    import chessboard_vision_b200.dropin as dropin
    golden = json.load(open(os.path.join(G, "callers.json")))
    monkeypatch.setitem(engine_mod._default, 0, fake)
    monkeypatch.syspath_prepend(dropin.PATH)
    mods = {}
    for name in ch.VISION:
        sys.modules.pop(name, None)
        mods[name] = importlib.import_module(name)
    try:
        g = ch.replay_game(mods, ch.scenario_game(), golden["game_session"])
        assert [[list(p) for p in o] for o in g["occupied"]] == golden["game_session"]["occupied"]
        assert [[list(p) for p in o] for o in g["changes"]] == golden["game_session"]["changes"]
        s = ch.replay_sensitivity(mods, ch.scenario_sensitivity(), golden["calibrate_sensitivity"]["loop_turns"])
        norm = lambda d: [{k: list(v) for k, v in x.items()} for x in d]
        assert norm(s["detailed"]) == norm(golden["calibrate_sensitivity"]["detailed"])
        assert s["changes"] == golden["calibrate_sensitivity"]["changes"]
    finally:
        for name in ch.VISION:
            sys.modules.pop(name, None)


def test_patched_draw_interface_in_a_live_session(fake):
    """INTEGRATION.md "Board overlay": GameSession._draw_interface replaced by BoardOverlay.session_state +
    draw_interface.  Inside the reference's own GameSession, on every frame of a game (grid from the calibration file,
    noise states, the e2e4 highlight, FPS text), the display list -- interpreted with OpenCV's arithmetic, no GPU here --
    is the picture the original method hands to cv2.imshow."""
    import numpy as np
    import cv2
    from chessboard_vision_b200.overlay import BoardOverlay
    frames = ch.scenario_game()
    with ch.caller_env("reference") as st:
        import game_session
        game_session.time = ch.FakeClock()
        original = game_session.GameSession._draw_interface
        shown, compared = [], []
        cv2.imshow = lambda name, img: shown.append((name, img.copy()))
        overlay = BoardOverlay(fake)

        def patched(self, vis, board_size, noise_state, img_raw):
            before = vis.copy()
            shown.clear()
            original(self, vis, board_size, noise_state, img_raw)
            want = [img for name, img in shown if name == "Tabuleiro"][0]
            with self.board_lock:
                state = BoardOverlay.session_state(self, noise_state == game_session.NoiseState.NOISE_ACTIVE)
            got = overlay.draw_interface(before, board_size, **state)
            compared.append((bool(np.array_equal(got, want)), bool(state["last_move"]), state["noise_active"]))
        game_session.GameSession._draw_interface = patched
        try:
            s = game_session.GameSession()
            cap = ch.FakeCap(frames)
            assert s.on_calibration_requested(cap)
            while True:
                ok, img = cap.read()
                if not ok:
                    break
                s.on_frame(img)
        finally:
            game_session.GameSession._draw_interface = original
    assert len(compared) >= 10 and all(c[0] for c in compared)
    assert any(c[1] for c in compared)                      # frames after the move carry the last-move highlight
