"""``BarotropicField`` — equivalent latitude and local wave activity of 2-D absolute-vorticity fields on the GPU, with
the interface of falwa.barotropic_field.BarotropicField (/root/reference/src/falwa/barotropic_field.py:13-164), plus
``BarotropicBatch`` for many time steps at once (BASELINE.json config 5)."""
from math import pi

import numpy as np
import torch

from . import _lib, engine
from ._lib import check, dptr, hptr, stream_ptr


class BarotropicBatch:
    """pv_fields [T][nlat][nlon] (NumPy or CUDA tensor) -> ``equivalent_latitudes`` [T][nlat], ``lwa`` [T][nlat][nlon]
    (or [T][2][nlat][nlon] = (anticyclonic, cyclonic) with ``return_partitioned_lwa``), device tensors."""

    def __init__(self, xlon, ylat, pv_fields, area=None, dphi=None, n_partitions=None, planet_radius=6.378e+6,
                 return_partitioned_lwa=False):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.xlon, self.ylat = np.asarray(xlon, float), np.asarray(ylat, float)
        self.nlon, self.nlat = self.xlon.size, self.ylat.size
        self.clat = np.abs(np.cos(np.deg2rad(self.ylat)))
        self.planet_radius = planet_radius
        self.return_partitioned_lwa = return_partitioned_lwa
        self.dphi = pi / (self.nlat - 1) * np.ones((self.nlat)) if dphi is None else np.asarray(dphi, float)
        if area is None:   # barotropic_field.py:84-91
            area = 2. * pi * planet_radius ** 2 * (np.cos(self.ylat[:, np.newaxis] * pi / 180.) *
                                                   self.dphi[:, np.newaxis]) / float(self.nlon) * \
                np.ones((self.nlat, self.nlon))
        self.area = np.ascontiguousarray(area, dtype=np.float64)
        self.n_partitions = self.nlat if n_partitions is None else int(n_partitions)
        if isinstance(pv_fields, torch.Tensor):
            assert pv_fields.is_cuda and pv_fields.dtype == torch.float64
            self.pv = pv_fields.contiguous()
        else:
            self.pv = torch.from_numpy(np.ascontiguousarray(pv_fields, dtype=np.float64)).cuda()
        if tuple(self.pv.shape[1:]) != (self.nlat, self.nlon):
            raise TypeError(f"Incorrect dimension of pv_field. Expected dimension: {(self.nlat, self.nlon)}")
        self._area_d = torch.from_numpy(self.area).cuda()
        self._eqvlat = self._lwa = None

    def compute(self, with_lwa=True):
        T = int(self.pv.shape[0])
        dmu = np.ascontiguousarray(self.planet_radius * self.clat * self.dphi)   # barotropic_field.py:121
        ylat = np.ascontiguousarray(self.ylat)
        qref = torch.empty((T, self.nlat), dtype=torch.float64, device=self.pv.device)
        lwa = None
        if with_lwa:
            shape = (T, 2, self.nlat, self.nlon) if self.return_partitioned_lwa else (T, self.nlat, self.nlon)
            lwa = torch.empty(shape, dtype=torch.float64, device=self.pv.device)
        check(self.lib.falwa_b200_barotropic_eqvlat_lwa(
            T, self.nlon, self.nlat, self.n_partitions, dptr(self.pv), dptr(self._area_d), hptr(ylat), hptr(dmu),
            float(self.planet_radius), dptr(qref), dptr(lwa), int(self.return_partitioned_lwa), stream_ptr()),
            "barotropic_eqvlat_lwa")
        self._eqvlat = qref
        if with_lwa:
            self._lwa = lwa
        return self

    @property
    def equivalent_latitudes(self):
        if self._eqvlat is None:
            self.compute(with_lwa=False)
        return self._eqvlat

    @property
    def lwa(self):
        if self._lwa is None:
            self.compute(with_lwa=True)
        return self._lwa


class BarotropicField:
    """Single 2-D field; same constructor and properties as the reference class."""

    def __init__(self, xlon, ylat, pv_field, area=None, dphi=None, n_partitions=None, planet_radius=6.378e+6,
                 return_partitioned_lwa=False):
        self._b = BarotropicBatch(xlon, ylat, np.asarray(pv_field)[None], area=area, dphi=dphi,
                                  n_partitions=n_partitions, planet_radius=planet_radius,
                                  return_partitioned_lwa=return_partitioned_lwa)
        for k in ("xlon", "ylat", "clat", "nlon", "nlat", "planet_radius", "return_partitioned_lwa", "dphi", "area",
                  "n_partitions"):
            setattr(self, k, getattr(self._b, k))
        self.pv_field = np.asarray(pv_field)

    @property
    def equivalent_latitudes(self):
        """Q(y) at the input latitudes, (nlat,)."""
        return engine.to_host(self._b.equivalent_latitudes[0])

    @property
    def lwa(self):
        """Local wave activity, (nlat, nlon) (or (2, nlat, nlon) when partitioned)."""
        return engine.to_host(self._b.lwa[0])
