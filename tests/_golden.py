import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_v1.json")))
ARR = np.load(os.path.join(HERE, "golden", "golden_v1_arrays.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_image(spec):
    from orbslam_in_practice_b200.synth import synth_frame, adversarial_frame
    if spec["kind"] == "synth":
        return synth_frame(spec["seed"], spec["w"], spec["h"])
    return adversarial_frame(spec["kind"], spec["w"], spec["h"])


def kp_sha(kps):
    if len(kps) == 0:
        return sha(np.zeros(0))
    return sha(np.stack([kps["x"], kps["y"], kps["size"], kps["response"], kps["octave"].astype(np.float32)], 1))
