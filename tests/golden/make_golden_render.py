"""Records whole images of the REFERENCE renderer (oracle/_ref/libref_render.so: the reference's own programs,
compiled unmodified, inside the emulated OptiX pipeline of oracle/ref_render.cpp) as tests/golden/
ref_render_vectors.json: per scene a CRC-32 of the float accumulation buffer's bit patterns plus a strided sample
of them.  Runs only where /root/reference exists; tests/test_reference_render.py replays the file against the
oracle anywhere.

    python tests/golden/make_golden_render.py
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as h  # noqa: E402
import test_reference_render as t  # noqa: E402


def main():
    if h.ReferenceRenderer.lib() is None:
        raise SystemExit("oracle/_ref/libref_render.so is not built (needs /root/reference)")
    cases = []
    for name, (build, W, H, frames, depth) in t.scene_table().items():
        ref = h.ReferenceRenderer(build())
        acc, _ = ref.render(W, H, 0, frames, max_depth=depth)
        assert np.isfinite(acc).all(), name
        bits = np.ascontiguousarray(acc, np.float32).view(np.uint32)
        stride = 97
        cases.append({"scene": name, "W": W, "H": H, "frames": frames, "max_depth": depth,
                      "crc32": zlib.crc32(bits.tobytes()), "stride": stride,
                      "bits": bits.reshape(-1)[::stride].tolist()})
    out = {"generator": "tests/golden/make_golden_render.py",
           "source": "reference programs compiled unmodified (oracle/ref_render.cpp), g++ argument order",
           "cases": cases}
    with open(os.path.join(HERE, "ref_render_vectors.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"{len(cases)} scenes recorded")


if __name__ == "__main__":
    main()
