"""Writes tests/golden/frame_stream.bin: the Frame stream of tests/test_frame_stream.py::make_fields(1) as written by the reference's
own Frame::toStream statements (oracle/_ref/libref_frame.so; needs /root/reference for its build: `make -C oracle ref`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from test_frame_stream import make_fields, ref_stream
ref_stream(make_fields(1)).tofile(os.path.join(ROOT, "tests", "golden", "frame_stream.bin"))
print("written")
from test_frame_stream import _mp_fields, _ref_mappoint
_ref_mappoint(_mp_fields(1)).tofile(os.path.join(ROOT, "tests", "golden", "mappoint_stream.bin"))
print("written mappoint")
from test_frame_stream import _ref_container
_ref_container([dict(_mp_fields(50 + i, n_frames=i), id=i) for i in range(6)], erase=[2]).tofile(os.path.join(ROOT, "tests", "golden", "mappoint_container.bin"))
print("written mappoint container")
# a complete small map FILE: the five sections of Map::toStream, each written by the reference's own code / statements, in the order of map.cpp:316-325
import tempfile, pathlib
import numpy as np
from test_frame_stream import _ref_sections, _ref_frame_container
S = _ref_sections(pathlib.Path(tempfile.mkdtemp()))
pts = _ref_container([dict(_mp_fields(70 + i, n_frames=2), id=i) for i in range(5)], erase=[3])
frs, _ = _ref_frame_container([dict(make_fields(40 + i, n_kp=50, n_markers=1), idx=i) for i in range(3)])
np.concatenate([np.frombuffer(np.uint64(225237123).tobytes(), np.uint8), S["kfdb"], pts, S["markers"], frs, S["covis"]]).tofile(
    os.path.join(ROOT, "tests", "golden", "map_file.bin"))
print("written map file")
