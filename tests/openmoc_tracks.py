"""Writer for OpenMOC track files in the layout the reference's reader consumes
(reference src/tracks.c:170-323, `-d <file>`): test infrastructure only.

    int string_length; char geometry[string_length];
    int n_azimuthal; double spacing;
    int num_tracks[n_azimuthal], num_x[n_azimuthal], num_y[n_azimuthal]; double azim_weights[n_azimuthal];
    per track:   double x0, y0, x1, y1, phi; int azim_angle_index; int num_segments;
      per segment: double length; int material_id; int region_id; (+ 2 ints of CMFD surfaces if cmfd)
"""
import struct

import numpy as np


def write_track_file(path, tracks_per_angle, segments, spacing=0.05, geometry="test geometry", cmfd=False):
    """tracks_per_angle: list of ints; segments: one list of (length, material_id, region_id) per track,
    azimuthal-angle major."""
    assert sum(tracks_per_angle) == len(segments)
    n_azim = len(tracks_per_angle)
    with open(path, "wb") as f:
        g = geometry.encode()
        f.write(struct.pack("=i", len(g)) + g)
        f.write(struct.pack("=id", n_azim, spacing))
        f.write(struct.pack(f"={n_azim}i", *tracks_per_angle))
        f.write(struct.pack(f"={n_azim}i", *[max(1, t // 2) for t in tracks_per_angle]))      # num_x
        f.write(struct.pack(f"={n_azim}i", *[t - max(1, t // 2) for t in tracks_per_angle]))  # num_y
        f.write(struct.pack(f"={n_azim}d", *[1.0 / n_azim] * n_azim))
        u = 0
        for a, count in enumerate(tracks_per_angle):
            phi = np.pi * (a + 0.5) / n_azim
            for j in range(count):
                segs = segments[u]
                f.write(struct.pack("=5dii", 0.1 * j, 0.0, 0.1 * j + np.cos(phi), np.sin(phi), phi, a, len(segs)))
                for length, material, region in segs:
                    f.write(struct.pack("=dii", length, material, region))
                    if cmfd:
                        f.write(struct.pack("=ii", -1, -1))
                u += 1
    return path


def synthetic_tracks(seed, tracks_per_angle, mean_segments, width=21.42, empty_every=0):
    """ragged random segmentation: every track crosses `width` in n ~ U[1, 2*mean) pieces; every
    `empty_every`-th track has no segment at all (the reference handles n_segments == 0)."""
    rng = np.random.default_rng(seed)
    out = []
    for u in range(sum(tracks_per_angle)):
        if empty_every and u % empty_every == empty_every - 1:
            out.append([])
            continue
        n = int(rng.integers(1, 2 * mean_segments))
        cuts = np.sort(rng.random(n - 1)) * width
        lengths = np.diff(np.concatenate(([0.0], cuts, [width])))
        out.append([(float(l), int(rng.integers(0, 7)), int(rng.integers(0, 1000))) for l in lengths])
    return out
