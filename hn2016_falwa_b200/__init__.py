"""hn2016_falwa_b200 — B200-native (sm_100a, fp64) implementation of falwa's QGField hot path.

Modules (same names and interfaces as in ``falwa`` where one exists):
  oopinterface      QGFieldNH18, QGFieldNHN22
  xarrayinterface   QGDataset, integrate_budget, hemisphere_to_globe (xarray objects or the Simple* stand-ins)
  barotropic_field  BarotropicField, BarotropicBatch
  basis, wrapper    the 2-D helper functions
  f90_modules       the nine F2PY routines by name (drop-in for falwa's native modules)
  netcdf_io         undecoded NetCDF-3 input, result files
  engine            batched device-resident pipeline (Plan, QGBatch, HostStream, run_stream, ingest_field)
"""
__version__ = "0.4.0"
