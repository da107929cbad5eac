"""CPU baseline of bench.py: processes x OpenMP threads sweep of the oracle pipeline at 0.25 degree on this box's host
cores (fp64), then the best split in fp32 ("as shipped": the reference's F2PY build is single precision,
src/falwa/f90_modules/meson.build:33-55) and one process x one thread.  Writes one JSON line per configuration."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    nlon, nlat = (1440, 721) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
    cores = os.cpu_count() or 1
    out = []
    splits = [w for w in (1, 2, 4, 8) if w <= cores]   # 16 x ~10 GB would not fit the host memory
    for w in splits:
        rate, c, workers, threads, wall = bench.cpu_reference_rate(nlon, nlat, snapshots=w, workers=w)
        rec = dict(grid=[nlon, nlat], precision="f64", processes=workers, threads=threads, snapshots=w,
                   snapshots_per_s=rate, wall_s=wall, cores=c)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    best = max(out, key=lambda r: r["snapshots_per_s"])
    rate, c, workers, threads, wall = bench.cpu_reference_rate(nlon, nlat, snapshots=best["processes"], workers=best["processes"],
                                                               precision="f32")
    print(json.dumps(dict(grid=[nlon, nlat], precision="f32", processes=workers, threads=threads, snapshots=best["processes"],
                          snapshots_per_s=rate, wall_s=wall, cores=c)), flush=True)
    rate, c, workers, threads, wall = bench.cpu_reference_rate(nlon, nlat, snapshots=1, workers=1, threads=1)
    print(json.dumps(dict(grid=[nlon, nlat], precision="f64", processes=1, threads=1, snapshots=1, snapshots_per_s=rate,
                          wall_s=wall, cores=c)), flush=True)


if __name__ == "__main__":   # the pool workers are spawned: they import this file
    main()
