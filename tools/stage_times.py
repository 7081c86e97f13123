"""Device time per 1080p frame of the stage kernels on their own (correct_lighting, bilateral, sharpen) beside the fused
process_pipeline path: what fusing saves and what each stage costs.    python tools/stage_times.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chessboard_vision_b200 import synth
from chessboard_vision_b200.engine import Engine
eng = Engine(0)
f = synth.frame_batch(8, 1080, 1920, "board", 0)
batch = np.stack([f[i % 8] for i in range(32)])
for name, fn in (("correct_lighting", lambda: eng.correct_lighting(batch)), ("bilateral", lambda: eng.bilateral(batch)),
                 ("sharpen", lambda: eng.sharpen(batch)), ("process_pipeline", lambda: eng.process_pipeline(batch))):
    fn()
    eng.profile(True); fn(); p = eng.profile_read(); eng.profile(False)
    print(name, {k: round(v[0] / v[1] / 32 * 1e3, 2) for k, v in p.items()}, "us per frame")
