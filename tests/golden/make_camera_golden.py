"""Generates tests/golden/camera_track.json: the camera pose after every frame of tests/golden/key_script.txt, computed
by the REFERENCE's own camera lines (renderer.cpp:438-467 compiled into oracle/_ref/libfs_camera_ref.so by
`make -C oracle ref`).  Needs /root/reference; run from the repo root."""
import ctypes as C
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fallingsand3d_b200.input import KeyFlags, KeyScript  # noqa: E402
from oracle import oracle  # noqa: E402

oracle.build_ref(force=True)
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfs_camera_ref.so"))
here = os.path.dirname(os.path.abspath(__file__))
text = open(os.path.join(here, "key_script.txt")).read()
script, flags = KeyScript(text), KeyFlags()
pos, rot = (C.c_float * 3)(), (C.c_float * 3)()
lib.fs_camera_ref_defaults(pos, rot)
frames = 160
track = []
for frame in range(frames):
    script.handle_events(frame, flags)
    lib.fs_camera_ref_step((C.c_uint8 * 8)(*[1 if f else 0 for f in flags.as_reference_order()]), pos, rot)
    track.append([struct.pack("<f", v).hex() for v in (pos[0], pos[1], pos[2], rot[1])])
with open(os.path.join(here, "camera_track.json"), "w") as f:
    json.dump({"source": "renderer.cpp:438-467 compiled with the reference's glm (oracle/_ref/libfs_camera_ref.so)",
               "frames": frames, "encoding": "little-endian float32 bytes as hex: x, y, z, yaw_deg", "track": track,
               "final": [pos[0], pos[1], pos[2], rot[1]]}, f, indent=0)
print("wrote camera_track.json; final pose", pos[0], pos[1], pos[2], rot[1])
