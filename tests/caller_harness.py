"""Drives the reference's UNCHANGED callers -- game_session.GameSession, play_lichess.main, calibrate_sensitivity.main
(SURVEY.md 8b last row) -- on synthetic camera frames, either on the reference's own vision modules or on the
drop-in modules of this repo (chessboard_vision_b200/dropin first on sys.path), and records what the callers saw:
occupancy per frame, visual changes, the move they inferred, the change dictionaries.

Needs the reference checkout (/root/reference): the callers are imported from there, never copied.  The GUI
(cv2.imshow / waitKey / namedWindow / trackbars: the headless OpenCV build raises), the camera (cv2.VideoCapture),
`input()` and the Lichess HTTP client are replaced by fakes; `chess` comes from tests/stubs.  Test infrastructure."""
import builtins
import contextlib
import os
import shutil
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = os.path.join(HERE, "stubs")
VISION = ("frame_enhancer", "grid_extractor", "board_detection", "change_detector", "piece_detector")
CALLERS = ("game_session", "lichess_session", "lichess_client", "play_lichess", "calibrate_sensitivity", "calibration_module",
           "game_state", "noise_handler", "chess")
GRID_X = [0, 79, 157, 234, 310, 386, 464, 541, 620]
GRID_Y = [0, 80, 158, 235, 311, 388, 465, 542, 620]
CORNERS = [[556, 112], [1560, 108], [1562, 1024], [550, 1005]]          # calibration.json:2-19 (a 1080p camera)


def have_reference():
    return os.path.isfile(os.path.join(REF, "game_session.py"))


def board_image(occupied):
    """A 620 x 620 top view: checker squares on the calibrated grid lines, a disc on every occupied (file, rank)."""
    img = np.zeros((620, 620, 3), np.uint8)
    yy, xx = np.ogrid[:620, :620]
    for r in range(8):
        for c in range(8):
            light = (r + c) % 2 == 0
            img[GRID_Y[r]:GRID_Y[r + 1], GRID_X[c]:GRID_X[c + 1]] = (150, 170, 180) if light else (70, 100, 120)
    for (f, rank) in occupied:
        r = 7 - rank
        cx, cy = (GRID_X[f] + GRID_X[f + 1]) // 2, (GRID_Y[r] + GRID_Y[r + 1]) // 2
        white = rank < 4
        img[(xx - cx) ** 2 + (yy - cy) ** 2 <= 29 * 29] = (250, 250, 252) if white else (15, 15, 20)
        img[(xx - cx) ** 2 + (yy - cy) ** 2 <= 9 * 9] = (215, 215, 220) if white else (50, 50, 55)
    return img


def camera_frames(positions, H=1080, W=1920, seed=0):
    """One 1080p camera frame per entry of `positions` (sets of occupied (file, rank)): the top view projected into
    the calibrated quadrilateral on a table-coloured background, plus a little sensor noise."""
    import cv2
    rng = np.random.default_rng(seed)
    tl, tr, br, bl = [np.float32(p) for p in CORNERS]
    M = cv2.getPerspectiveTransform(np.float32([tl, tr, bl, br]), np.float32([[0, 0], [620, 0], [0, 620], [620, 620]]))
    cache, out = {}, []
    for occ in positions:
        key = frozenset(occ)
        if key not in cache:
            cache[key] = cv2.warpPerspective(board_image(occ), M, (W, H), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                                             borderMode=cv2.BORDER_CONSTANT, borderValue=(95, 115, 125))
        f = cache[key].astype(np.int16) + rng.integers(-2, 3, (H, W, 3))
        out.append(np.clip(f, 0, 255).astype(np.uint8))
    return out


START = {(f, r) for f in range(8) for r in (0, 1, 6, 7)}
AFTER_E2E4 = (START - {(4, 1)}) | {(4, 3)}


class FakeCap:
    def __init__(self, frames):
        self.frames, self.i = frames, 0

    def isOpened(self):
        return True

    def set(self, *a):
        return True

    def read(self):
        if self.i >= len(self.frames):
            return False, None
        self.i += 1
        return True, self.frames[self.i - 1].copy()

    def release(self):
        pass


class FakeClock:
    """Stands in for the `time` module inside game_session: 20 ms per call, so that the once-per-second full scan of
    on_frame (game_session.py:115-121,137) falls on the same frames in every run."""

    def __init__(self):
        self.t = 1000.0

    def time(self):
        self.t += 0.02
        return self.t

    def sleep(self, s):
        self.t += s


class MockLichessClient:
    """lichess_client.LichessClient without the network (lichess_session.py:12,24,53,69,110-124)."""
    my_color = "white"
    sent = []

    def connect(self):
        return True

    def get_ongoing_games(self):
        return [{"gameId": "smoke01", "opponent": {"username": "nobody"}}]

    def stream_game(self, game_id):
        return iter([{"type": "gameFull", "state": {"moves": ""}}])

    def make_move(self, uci):
        MockLichessClient.sent.append(uci)
        return True

    def is_my_turn(self, moves):
        return len(moves.split()) % 2 == 0

    def get_last_move(self, moves):
        return moves.split()[-1] if moves else None


@contextlib.contextmanager
def caller_env(kind, engine=None, answers=("s", "1"), quit_after=10 ** 9, trackbars=None):
    """sys.path / cwd / GUI / camera patched for one run.  kind: 'reference' (the reference's own vision modules) or
    'dropin' (this repo's modules first on sys.path, `engine` installed as the default engine of device 0)."""
    import cv2
    import chessboard_vision_b200.dropin as dropin
    import chessboard_vision_b200.engine as engine_mod
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    saved_mods = {m: sys.modules.pop(m) for m in list(sys.modules) if m in VISION + CALLERS or m.startswith("_cvb_ref_")}
    saved_default = dict(engine_mod._default)
    saved_cv = {n: getattr(cv2, n) for n in ("imshow", "waitKey", "namedWindow", "destroyAllWindows", "setMouseCallback",
                                             "createTrackbar", "getTrackbarPos", "resizeWindow", "VideoCapture", "moveWindow")
                if hasattr(cv2, n)}
    saved_input = builtins.input
    tmp = tempfile.mkdtemp(prefix="callers_cwd_")
    shutil.copy(os.path.join(REF, "calibration.json"), tmp)       # CalibrationModule.run offers to load it (answer 's')
    state = {"keys": 0, "frames": None, "answers": list(answers), "bars": dict(trackbars or {}), "shown": 0}
    try:
        sys.path[:0] = [STUBS] + ([dropin.PATH] if kind == "dropin" else []) + [REF]
        if kind == "dropin":
            engine_mod._default[0] = engine
        os.chdir(tmp)

        def wait_key(_=0):
            state["keys"] += 1
            return ord("q") if state["keys"] >= quit_after else -1
        cv2.imshow = lambda *a: state.__setitem__("shown", state["shown"] + 1)
        cv2.waitKey = wait_key
        for n in ("namedWindow", "destroyAllWindows", "setMouseCallback", "resizeWindow", "moveWindow"):
            setattr(cv2, n, lambda *a, **k: None)
        cv2.createTrackbar = lambda name, win, val, mx, cb: state["bars"].__setitem__(name, val)
        cv2.getTrackbarPos = lambda name, win: state["bars"][name]
        cv2.VideoCapture = lambda *a: FakeCap(state["frames"])
        builtins.input = lambda *a: state["answers"].pop(0) if state["answers"] else "n"
        yield state
    finally:
        builtins.input = saved_input
        for n, v in saved_cv.items():
            setattr(cv2, n, v)
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for m in list(sys.modules):
            if m in VISION + CALLERS or m.startswith("_cvb_ref_"):
                del sys.modules[m]
        sys.modules.update(saved_mods)
        engine_mod._default.clear(); engine_mod._default.update(saved_default)
        shutil.rmtree(tmp, ignore_errors=True)


def _which(mod):
    return "dropin" if "chessboard_vision_b200" in (getattr(mod, "__file__", "") or "") else "reference"


def run_game_session(kind, frames, engine=None):
    """GameSession.on_calibration_requested + on_frame over `frames` (game_session.py:57-179) -> record."""
    with caller_env(kind, engine) as st:
        import game_session
        assert game_session.__file__.startswith(REF)
        game_session.time = FakeClock()
        rec = {"vision": _which(sys.modules["piece_detector"]), "occupied": [], "changes": [], "noise": [], "moves": [],
               "to_check": [], "updates": []}
        s = game_session.GameSession()
        cap = FakeCap(frames)
        assert s.on_calibration_requested(cap)
        inner, detect = s._process_stable_move, s.piece_detector.detect_all_pieces

        def spy_detect(squares, use_smoothing=True, use_delta=True, squares_to_check=None):
            res, vis = detect(squares, use_smoothing=use_smoothing, use_delta=use_delta, squares_to_check=squares_to_check)
            rec["changes"].append(sorted(vis))
            rec["to_check"].append(None if squares_to_check is None else sorted(squares_to_check))
            return res, vis
        update = s.piece_detector.update_references

        def spy_update(squares):
            rec["updates"].append(len(rec["occupied"]) - 1)      # index of the frame whose squares become the references
            return update(squares)
        s.piece_detector.update_references = spy_update

        def spy_move(vision_occupied, squares, noise_state):
            rec["occupied"].append(sorted(vision_occupied)); rec["noise"].append(noise_state.name)
            return inner(vision_occupied, squares, noise_state)
        s.piece_detector.detect_all_pieces, s._process_stable_move = spy_detect, spy_move
        while True:
            ok, img = cap.read()
            if not ok:
                break
            s.on_frame(img)
        rec["moves"] = [m.uci() for m in s.game.board.move_stack]
        rec["status"], rec["frames"], rec["shown"] = s.status, len(rec["occupied"]), st["shown"]
        return rec


def run_play_lichess(kind, frames, engine=None, loop_turns=12):
    """play_lichess.main() (play_lichess.py:14-75) with a fake camera and Lichess client -> record."""
    with caller_env(kind, engine, quit_after=loop_turns) as st:
        st["frames"] = frames
        import lichess_session
        import game_session
        game_session.time = FakeClock()
        lichess_session.LichessClient = MockLichessClient
        MockLichessClient.sent = []
        import play_lichess
        calls = []
        orig = lichess_session.LichessSession.on_frame
        lichess_session.LichessSession.on_frame = lambda self, img: (calls.append(img.shape), orig(self, img))[1]
        play_lichess.main()
        return {"vision": _which(sys.modules["piece_detector"]), "on_frame_calls": len(calls), "key_polls": st["keys"],
                "shown": st["shown"], "sent": list(MockLichessClient.sent)}


def run_calibrate_sensitivity(kind, frames, engine=None, loop_turns=40):
    """calibrate_sensitivity.main() (calibrate_sensitivity.py:62-394; the loop of :110-162) -> the dictionaries
    ChangeDetector returned on every frame after the automatic calibration at frame 30."""
    with caller_env(kind, engine, answers=("s",), quit_after=loop_turns) as st:
        st["frames"] = frames
        import calibrate_sensitivity as cs
        import change_detector
        rec = {"vision": _which(change_detector), "detailed": [], "changes": [], "patterns": []}
        cls = change_detector.ChangeDetector
        d0, c0, h0 = cls.detect_changes_detailed, cls.detect_changes, cls.classify_hand_pattern

        nested = []

        def spy_detailed(self, squares):
            out = d0(self, squares)
            if nested:                      # called from inside detect_changes (change_detector.py:96): not a caller's call
                return out
            rec["detailed"].append({"%d_%d" % p: (v["intensity"], round(v["pct_changed"], 6), bool(v["is_circular"]))
                                    for p, v in sorted(out.items())})
            return out

        def spy_changes(self, squares):
            nested.append(1)
            try:
                out = c0(self, squares)
            finally:
                nested.pop()
            rec["changes"].append(sorted("%d_%d" % p for p in out))
            return out

        def spy_pattern(self, detailed):
            out = h0(self, detailed)
            rec["patterns"].append((bool(out["is_hand"]), bool(out["is_move"]), sorted(out["move_candidates"])))
            return out
        cls.detect_changes_detailed, cls.detect_changes, cls.classify_hand_pattern = spy_detailed, spy_changes, spy_pattern
        try:
            cs.main()
        finally:
            cls.detect_changes_detailed, cls.detect_changes, cls.classify_hand_pattern = d0, c0, h0
        rec["shown"] = st["shown"]
        return rec


# ---- scenarios shared by the tests and by tools/make_golden_callers.py ----
def scenario_game():
    return camera_frames([START] * 19 + [AFTER_E2E4] * 30, seed=0)


def scenario_sensitivity():
    hand = START | {(3, 3), (3, 4), (4, 4), (2, 3)}
    return camera_frames([START] * 33 + [AFTER_E2E4] * 4 + [hand] * 3, seed=1)


def replay_game(mods, frames, golden):
    """The vision calls GameSession makes (game_session.py:57-111 capture_reference, :113-161 on_frame), restated for
    boxes without the reference checkout, with the `squares_to_check` sets the golden run passed -> record."""
    import cv2
    bd, ge, pdm = mods["board_detection"], mods["grid_extractor"], mods["piece_detector"]
    pts = bd.reorder(np.array(CORNERS).reshape((4, 1, 2)))
    grid = ge.SmartGridExtractor()
    grid.grid_lines_x, grid.grid_lines_y = list(GRID_X), list(GRID_Y)
    pd = pdm.PieceDetector()
    warped, _, _ = bd.warp_image(frames[10], pts)                 # ten reads to settle, the eleventh is the reference
    pd.update_references(grid.split_board(warped))
    rec = {"occupied": [], "changes": []}
    for i, img in enumerate(frames[11:]):
        warped, _, board_size = bd.warp_image(img, pts)
        squares = grid.split_board(warped)
        tc = golden["to_check"][i]
        res, vis = pd.detect_all_pieces(squares, use_delta=True, squares_to_check=None if tc is None else {tuple(p) for p in tc})
        rec["occupied"].append(sorted(p for p, info in res.items() if info["has_piece"]))
        rec["changes"].append(sorted(vis))
        if i in golden["updates"]:
            pd.update_references(squares)
    return rec


def replay_sensitivity(mods, frames, n_turns):
    """The loop body of calibrate_sensitivity.py:110-162 with its default trackbar values, restated."""
    bd, ge, cdm = mods["board_detection"], mods["grid_extractor"], mods["change_detector"]
    pts = bd.reorder(np.array(CORNERS).reshape((4, 1, 2)))
    grid, det = ge.GridExtractor(), cdm.ChangeDetector()
    rec = {"detailed": [], "changes": [], "patterns": []}
    for fc, img in enumerate(frames[:n_turns]):
        det.z_threshold, det.initial_variance, det.alpha, det.blur_kernel = 2.0, 100, 0.2, 5   # DEFAULT_SETTINGS through the trackbars
        det._kernel = 5
        warped, _, _ = bd.warp_image(img, pts)
        squares = grid.split_board(warped)
        if fc == 30 and not det.is_calibrated:
            det.calibrate(squares)
        if not det.is_calibrated:
            continue
        detailed = det.detect_changes_detailed(squares)
        rec["detailed"].append({"%d_%d" % p: (v["intensity"], round(v["pct_changed"], 6), bool(v["is_circular"]))
                                for p, v in sorted(detailed.items())})
        rec["changes"].append(sorted("%d_%d" % p for p in det.detect_changes(squares)))
        if detailed:
            out = det.classify_hand_pattern(detailed)
            rec["patterns"].append((bool(out["is_hand"]), bool(out["is_move"]), sorted(out["move_candidates"])))
    return rec
