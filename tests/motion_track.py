"""TEST INFRASTRUCTURE: deterministic receiver track used by the motion scenarios."""


def write_motion(path, npts):
    """Deterministic 10 Hz ECEF track: a 150 m circle flown at 15 m/s around the static location,
    climbing 0.5 m/s (stand-in for the reference's circle.csv, which cannot travel to the GPU box)."""
    import math
    lat, lon, h = math.radians(35.681298), math.radians(139.766247), 10.0
    a, e2 = 6378137.0, 0.0818191908426 ** 2
    n = a / math.sqrt(1 - e2 * math.sin(lat) ** 2)
    x0 = (n + h) * math.cos(lat) * math.cos(lon)
    y0 = (n + h) * math.cos(lat) * math.sin(lon)
    z0 = ((1 - e2) * n + h) * math.sin(lat)
    north = (-math.sin(lat) * math.cos(lon), -math.sin(lat) * math.sin(lon), math.cos(lat))
    east = (-math.sin(lon), math.cos(lon), 0.0)
    up = (math.cos(lat) * math.cos(lon), math.cos(lat) * math.sin(lon), math.sin(lat))
    with open(path, "w") as f:
        for i in range(npts):
            t = 0.1 * i
            ang = 15.0 * t / 150.0
            dn, de, du = 150.0 * math.sin(ang), 150.0 * (1 - math.cos(ang)), 0.5 * t
            p = [c0 + dn * nn + de * ee + du * uu for c0, nn, ee, uu in zip((x0, y0, z0), north, east, up)]
            f.write("%5.1f,%.3f,%.3f,%.3f\n" % (t, p[0], p[1], p[2]))
