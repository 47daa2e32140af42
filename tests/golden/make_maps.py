"""Generate tests/golden/maps/*.npz from the reference's PGM/YAML fixtures (mpros_planner/map/).

Run HERE (needs /root/reference); the .npz files are committed so nothing reads the reference at
test time. Conversion follows ROS map_server's trinary mode (third-party, absent):
p = (255 - px) / 255 (negate: 0); p > occupied_thresh -> 100; p < free_thresh -> 0; else -1;
image row 0 is the TOP, grid row 0 the BOTTOM.
"""
import os
import re
import sys

import numpy as np

REF_MAPS = "/root/reference/mpros_planner/map"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maps")
NAMES = ["reverse_parking", "parallel_parking3", "parking4", "ARENA_part", "map_slalom", "parkinglot"]


def read_pgm(path):
    with open(path, "rb") as f:
        buf = f.read()
    # header: magic, width, height, maxval separated by whitespace, '#' comments allowed
    tokens, pos = [], 0
    while len(tokens) < 4:
        m = re.compile(rb"\s*(#[^\n]*\n\s*)*(\S+)").match(buf, pos)
        tokens.append(m.group(2))
        pos = m.end()
    magic, w, h, maxval = tokens[0], int(tokens[1]), int(tokens[2]), int(tokens[3])
    if magic == b"P5":
        pos += 1  # single whitespace after maxval
        dt = np.uint8 if maxval < 256 else np.dtype(">u2")
        img = np.frombuffer(buf, dtype=dt, count=w * h, offset=pos).reshape(h, w)
    elif magic == b"P2":
        img = np.array(buf[pos:].split(), dtype=np.int64).reshape(h, w)
    else:
        raise ValueError("unsupported PGM magic %r" % magic)
    return img.astype(np.float64) * (255.0 / maxval)


def read_yaml(path):
    out = {}
    for line in open(path):
        line = line.split("#")[0].strip()
        if ":" not in line:
            continue
        k, v = line.split(":", 1)
        out[k.strip()] = v.strip()
    out["resolution"] = float(out["resolution"])
    out["origin"] = [float(x) for x in out["origin"].strip("[]").split(",")]
    out["negate"] = int(out["negate"])
    out["occupied_thresh"] = float(out["occupied_thresh"])
    out["free_thresh"] = float(out["free_thresh"])
    return out


def to_occupancy(img, meta):
    p = img / 255.0 if meta["negate"] else (255.0 - img) / 255.0
    occ = np.full(img.shape, -1, dtype=np.int8)
    occ[p > meta["occupied_thresh"]] = 100
    occ[p < meta["free_thresh"]] = 0
    return np.ascontiguousarray(occ[::-1])  # flip: grid row 0 = bottom image row


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in NAMES:
        meta = read_yaml(os.path.join(REF_MAPS, name + ".yaml"))
        img = read_pgm(os.path.join(REF_MAPS, os.path.basename(meta["image"])))
        occ = to_occupancy(img, meta)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            occupancy=occ,
            resolution=np.float64(meta["resolution"]),
            origin=np.array(meta["origin"], dtype=np.float64),
            occupied_thresh=np.float64(meta["occupied_thresh"]),
            free_thresh=np.float64(meta["free_thresh"]),
        )
        vals, cnt = np.unique(occ, return_counts=True)
        print(name, occ.shape, meta["resolution"], meta["origin"], dict(zip(vals.tolist(), cnt.tolist())))
    # case 3 stand-in: Arena2036.pgm is missing from the reference, so ARENA_part's image is read with
    # Arena2036.yaml's geometry and thresholds (res 0.04, origin 0, occupied 0.99, free 0.6); the tests scale it x4
    # (nearest neighbour) to 2000 x 2000
    meta = read_yaml(os.path.join(REF_MAPS, "Arena2036.yaml"))
    img = read_pgm(os.path.join(REF_MAPS, "ARENA_part.pgm"))
    occ = to_occupancy(img, meta)
    np.savez_compressed(os.path.join(OUT, "arena_standin.npz"), occupancy=occ, resolution=np.float64(meta["resolution"]),
                        origin=np.array(meta["origin"], dtype=np.float64), occupied_thresh=np.float64(meta["occupied_thresh"]),
                        free_thresh=np.float64(meta["free_thresh"]))
    vals, cnt = np.unique(occ, return_counts=True)
    print("arena_standin", occ.shape, meta["resolution"], meta["origin"], dict(zip(vals.tolist(), cnt.tolist())))


if __name__ == "__main__":
    sys.exit(main())
