"""Key flags, camera integration and paint brush of the frame loop — the Python mirror of include/fs3d_input.hpp.

Reference: the eight key flags of engine::Window (/root/reference/src/engine/window.h:12-19, set by
window.cpp:34-107) and the camera integration at the top of Renderer::draw
(/root/reference/src/engine/rendering/renderer.cpp:438-467: 1.5 units/s, 10 degrees/s, fixed dt 0.016, defaults
renderer.h:148-149).  All camera arithmetic is float32, in the reference's order, so a replayed key script gives
bit-identical floats to the C++ header and to the reference's own lines compiled in oracle/_ref.
The paint / erase flags, the material selector and the brush are builder-defined (the reference has no paint input).
"""
import math

import numpy as np

_F = np.float32
KEYS = ("W", "A", "S", "D", "LCTRL", "SPACE", "LEFT", "RIGHT")


class KeyFlags:
    def __init__(self):
        self.holdingW = self.holdingA = self.holdingS = self.holdingD = False
        self.holdingCTRL = self.holdingSpace = self.holdingLeft = self.holdingRight = False
        self.holdingPaint = self.holdingErase = False
        self.material = 1

    def on_key(self, key, down):
        names = {"W": "holdingW", "A": "holdingA", "S": "holdingS", "D": "holdingD", "LCTRL": "holdingCTRL",
                 "SPACE": "holdingSpace", "LEFT": "holdingLeft", "RIGHT": "holdingRight", "PAINT": "holdingPaint",
                 "ERASE": "holdingErase"}
        if key in names:
            setattr(self, names[key], bool(down))
        elif down and len(key) == 1 and "1" <= key <= "7":
            self.material = int(key)

    def as_reference_order(self):
        """W, A, S, D, CTRL, Space, Left, Right — the order of window.h:12-19."""
        return [self.holdingW, self.holdingA, self.holdingS, self.holdingD, self.holdingCTRL, self.holdingSpace,
                self.holdingLeft, self.holdingRight]


class CameraController:
    def __init__(self):
        self.cam_pos = np.array([0.0, 0.0, -5.0], dtype=_F)    # renderer.h:148
        self.cam_rot = np.array([0.0, 0.0, 0.0], dtype=_F)     # renderer.h:149
        self.cam_move_speed = _F(1.5)
        self.cam_rot_speed = _F(10.0)

    def integrate(self, k, dt=0.016):
        dt = _F(dt)
        mv, rt = _F(self.cam_move_speed * dt), _F(self.cam_rot_speed * dt)
        p, r = self.cam_pos, self.cam_rot
        if k.holdingW:
            p[2] = _F(p[2] + mv)
        elif k.holdingS:
            p[2] = _F(p[2] - mv)
        if k.holdingA:
            p[0] = _F(p[0] + mv)
        elif k.holdingD:
            p[0] = _F(p[0] - mv)
        if k.holdingSpace:
            p[1] = _F(p[1] - mv)
        elif k.holdingCTRL:
            p[1] = _F(p[1] + mv)
        if k.holdingRight:
            r[1] = _F(r[1] + rt)
        elif k.holdingLeft:
            r[1] = _F(r[1] - rt)

    def camera(self, aspect=1700.0 / 900.0):
        """keyword arguments for VoxelWorld.raymarch"""
        return dict(pos=tuple(float(v) for v in self.cam_pos), yaw_deg=float(self.cam_rot[1]), aspect=aspect)


def brush_centre(cam, distance, nx, ny, nz):
    yaw = float(cam.cam_rot[1]) * 3.14159265358979323846 / 180.0
    px = float(cam.cam_pos[0]) + distance * math.sin(yaw)
    py = float(cam.cam_pos[1])
    pz = float(cam.cam_pos[2]) + distance * math.cos(yaw)
    h = 1.0 / max(nx, ny, nz)
    return (int(math.floor((px + 0.5 * nx * h) / h)), ny - 1 - int(math.floor((py + 0.5 * ny * h) / h)),
            int(math.floor((pz + 0.5 * nz * h) / h)))


class KeyScript:
    """'frame key down|up' per line; '#' starts a comment."""

    def __init__(self, text):
        self.events = []
        for line in text.splitlines():
            line = line.split("#", 1)[0].split()
            if len(line) >= 3:
                self.events.append((int(line[0]), line[1], line[2] == "down"))
        self._next = 0

    def handle_events(self, frame, flags):
        while self._next < len(self.events) and self.events[self._next][0] <= frame:
            _, key, down = self.events[self._next]
            flags.on_key(key, down)
            self._next += 1

    @property
    def last_frame(self):
        return self.events[-1][0] if self.events else 0


def replay(world, script_text, frames, steps_per_frame=1, brush_distance=4.5, brush_radius=3, on_frame=None):
    """The headless frame loop (engine.cpp:59-70): handleEvents -> camera integration (+ brush) -> step.
    Returns the camera track [(x, y, z, yaw)] after every frame."""
    script, flags, cam = KeyScript(script_text), KeyFlags(), CameraController()
    track = []
    for frame in range(frames):
        script.handle_events(frame, flags)
        cam.integrate(flags)
        if world is not None:
            if flags.holdingPaint or flags.holdingErase:
                c = brush_centre(cam, brush_distance, world.nx, world.ny, world.nz)
                world.paint_sphere(c, brush_radius, 0 if flags.holdingErase else flags.material,
                                   only_empty=not flags.holdingErase)
            world.step(steps_per_frame)
        track.append((float(cam.cam_pos[0]), float(cam.cam_pos[1]), float(cam.cam_pos[2]), float(cam.cam_rot[1])))
        if on_frame is not None:
            on_frame(frame, cam, flags)
    return track
