#!/usr/bin/env python
"""One file of 6-hourly u, v, t on pressure levels -> LWA budget terms on the GPU -> NetCDF, the workflow of the
reference's production script (scripts/lubis_et_al_2025/cal_lwa_budget_NHN22.py) with hn2016_falwa_b200.

    python examples/lwa_budget_nhn22.py u.nc v.nc t.nc output.nc            # NetCDF-3 inputs (packed int16 or float)
    torchrun --nproc-per-node 8 examples/lwa_budget_nhn22.py ... out.nc     # time-sharded over 8 GPUs

The packed values are shipped to the GPU as they lie in the files (2 bytes per value) and decoded, byte-swapped and
reordered there; every time step runs through QGFieldNHN22's three stages in device batches; the per-rank outputs are
gathered once at the end (the only communication of the job)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hn2016_falwa_b200 import netcdf_io                                     # noqa: E402
from hn2016_falwa_b200.oopinterface import QGFieldNHN22                     # noqa: E402
from hn2016_falwa_b200.xarrayinterface import QGDataset                     # noqa: E402


def main(u_path, v_path, t_path, out_path):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl")
    ds = netcdf_io.open_fields(u_path, v_path, t_path)
    q = QGDataset(ds, qgfield=QGFieldNHN22, keep_3d=False,
                  qgfield_kwargs=dict(kmax=49, dz=1000., eq_boundary_index=5, northern_hemisphere_results_only=False))
    q.compute_lwa_and_barotropic_fluxes(return_dataset=False)
    q.gather()
    if world == 1 or torch.distributed.get_rank() == 0:
        names = ["lwa_baro", "u_baro", "adv_flux_f1", "adv_flux_f2", "adv_flux_f3", "convergence_zonal_advective_flux",
                 "divergence_eddy_momentum_flux", "meridional_heat_flux", "qref", "uref", "ptref"]
        netcdf_io.write_results(out_path, q._dataset(names), names=names)
        print(f"wrote {len(names)} variables for {q._n} time steps to {out_path}; "
              f"mean barotropic LWA {float(np.mean(q.lwa_baro.data)):.3f} m/s")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    if len(sys.argv) != 5:
        sys.exit(__doc__)
    main(*sys.argv[1:])
