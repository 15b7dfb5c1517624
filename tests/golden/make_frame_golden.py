"""Writes tests/golden/frame_stream.bin: the Frame stream of tests/test_frame_stream.py::make_fields(1) as written by the reference's
own Frame::toStream statements (oracle/_ref/libref_frame.so; needs /root/reference for its build: `make -C oracle ref`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_frame_stream import make_fields, ref_stream
ref_stream(make_fields(1)).tofile(os.path.join(ROOT, "tests", "golden", "frame_stream.bin"))
print("written")
from test_frame_stream import _mp_fields, _ref_mappoint
_ref_mappoint(_mp_fields(1)).tofile(os.path.join(ROOT, "tests", "golden", "mappoint_stream.bin"))
print("written mappoint")
from test_frame_stream import _ref_container
_ref_container([dict(_mp_fields(50 + i, n_frames=i), id=i) for i in range(6)], erase=[2]).tofile(os.path.join(ROOT, "tests", "golden", "mappoint_container.bin"))
print("written mappoint container")
