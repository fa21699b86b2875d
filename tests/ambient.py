"""Host-side construction of AmbientHeating's static heating plane (reference source/modules/solar/ambientheating.cpp:28-40),
done with the host libm exactly once at setup, as the reference does; the device only adds dt*heating."""
import numpy as np


def heating_plane(g, kw):
    nx, ny = g.planes["rho"].shape
    xb, yb = g.kw["xb"], g.kw["yb"]
    mask = np.zeros((nx, ny))
    xl = 0 if xb[0] == "periodic" else 2
    xu = nx - 1 if xb[1] == "periodic" else nx - 3
    yl = 0 if yb[0] == "periodic" else 2
    yu = ny - 1 if yb[1] == "periodic" else ny - 3
    mask[xl:xu + 1, yl:yu + 1] = 1.0
    if not kw["exp_mode"]:
        return mask * kw["heating_rate"]
    import math
    pos_y = g.planes["pos_y"]
    ex = np.vectorize(math.exp)
    h = (mask * kw["exp_base_heating_rate"]) * ex((pos_y * -1.0) / kw["exp_scale_height"])
    if kw["split_exp_mode"]:
        H, Hs, y0 = kw["exp_scale_height"], kw["split_exp_scale_height"], kw["split_exp_start_height"]
        sb = kw["exp_base_heating_rate"] * math.exp((H - Hs) * y0 / (H * Hs))
        h = np.maximum(h, (mask * sb) * ex((pos_y * -1.0) / Hs))
    return h
