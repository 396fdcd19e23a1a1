"""Generates tests/golden/tracks_44.bin, the OpenMOC-format track file the loader tests read
(44 2D tracks over 4 azimuthal angles, 1-23 segments each, every 9th track empty):

    python tests/golden/make_track_file.py

The file is committed (numpy's Generator stream is not guaranteed across versions)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from openmoc_tracks import synthetic_tracks, write_track_file  # noqa: E402

PER_ANGLE = [12, 10, 12, 10]

if __name__ == "__main__":
    segs = synthetic_tracks(20261017, PER_ANGLE, 12, empty_every=9)
    p = write_track_file(os.path.join(HERE, "tracks_44.bin"), PER_ANGLE, segs, spacing=0.0371)
    print(p, os.path.getsize(p), "bytes,", sum(len(s) for s in segs), "segments")
