"""Trajectory output (reference: pyparm/xyzfile.py:7-76 XYZwriter, src/bin/LJatoms.cpp:130-158 writefile).

Frames come from parm_snapshot_begin / parm_snapshot_wait: the positions and velocities are gathered on the device
and copied to page-locked host memory on a second stream, so the run keeps stepping while a frame travels and is
formatted. File layout as in the reference: atom count, a comment line of key=value pairs (`time` is mandatory), then
one line per atom: element, x y z (%.4f), optionally vx vy vz."""
import numpy as np


class XYZwriter:
    def __init__(self, f, usevels=True, elements=None):
        self.file = f
        self.usevels = usevels
        self.elements = elements   # per-atom element symbols (default "C" as LJatoms.cpp:143)
        self._pending = None

    # -- asynchronous interface -------------------------------------------------------------------------------------
    def begin_frame(self, atoms, com=None, box=None, **kwargs):
        """Start the device->host copy of this frame and return; finish_frame() (called automatically by the next
        begin_frame, writeframe or close) formats and writes it."""
        if "time" not in kwargs:
            raise ValueError("Time must be in keyword arguments, or .xyz will not be readable")
        self.finish_frame()
        atoms.snapshot_begin(velocities=self.usevels)
        self._pending = (atoms, None if com is None else np.asarray(com, dtype=np.float64), box, kwargs)

    def finish_frame(self):
        if self._pending is None:
            return
        atoms, com, box, kwargs = self._pending
        self._pending = None
        x, v = atoms.snapshot_wait()
        if com is not None or box is not None:
            ref = np.zeros(x.shape[1]) if com is None else com
            x = x - ref if box is None else box.diff(x, np.broadcast_to(ref, x.shape).copy(), atoms)
        print(len(x), file=self.file)
        print(" ".join("=".join((k, str(val))) for k, val in kwargs.items()), file=self.file)
        el = self.elements if self.elements is not None else ["C"] * len(x)
        cols = [x] + ([v] if self.usevels and v is not None else [])
        data = np.concatenate(cols, axis=1)
        fmt = " ".join(["%.4f"] * data.shape[1])
        self.file.write("\n".join(e + " " + fmt % tuple(row) for e, row in zip(el, data)))
        self.file.write("\n")
        self.file.flush()

    # -- the reference's interface ----------------------------------------------------------------------------------
    def writeframe(self, atoms, com=None, box=None, **kwargs):
        self.begin_frame(atoms, com, box, **kwargs)
        self.finish_frame()

    def writefull(self, t, atoms, collec, com=None):
        cdict = {"time": t, "E": collec.energy(), "T": collec.temp(), "K": collec.kinetic_energy(),
                 "v": float(np.linalg.norm(collec.com_velocity()))}
        if com is None:
            com = atoms.com()
        self.writeframe(atoms, com, **cdict)
        return cdict

    def size(self):
        location = self.file.tell()
        self.file.seek(0, 2)
        size = self.file.tell()
        self.file.seek(location)
        return size

    def close(self):
        self.finish_frame()
        self.file.close()


class NPZwriter:
    """Frames accumulated from asynchronous downloads and saved as one .npz (times, x[frame, atom, dim], v)."""

    def __init__(self, path, usevels=True):
        self.path, self.usevels = path, usevels
        self.t, self.x, self.v = [], [], []
        self._pending = None

    def begin_frame(self, atoms, time):
        self.finish_frame()
        atoms.snapshot_begin(velocities=self.usevels)
        self._pending = (atoms, time)

    def finish_frame(self):
        if self._pending is None:
            return
        atoms, time = self._pending
        self._pending = None
        x, v = atoms.snapshot_wait()
        self.t.append(time)
        self.x.append(x)
        if v is not None:
            self.v.append(v)

    def close(self):
        self.finish_frame()
        out = {"time": np.asarray(self.t), "x": np.asarray(self.x)}
        if self.v:
            out["v"] = np.asarray(self.v)
        np.savez(self.path, **out)
