"""``QGFieldNH18`` / ``QGFieldNHN22`` — same constructor, methods, properties, named tuples and error behaviour
as ``falwa.oopinterface`` (/root/reference/src/falwa/oopinterface.py:26-1927), computed on one B200 by the
kernels of libfalwa_b200.so.  All fields stay on the device between stages; properties copy to NumPy lazily.

Differences from the reference that are visible to a user:
  * float64 throughout (the shipped F2PY build is float32 — SURVEY F1);
  * regions the reference leaves uninitialised are 0;
  * no printing (the reference prints from Python and Fortran — SURVEY §5).
"""
import warnings
from abc import ABC, abstractmethod
from collections import namedtuple
from typing import NamedTuple, Optional

import numpy as np
from scipy.interpolate import interp1d

from . import engine
from .constant import P_GROUND, SCALE_HEIGHT, CP, DRY_GAS_CONSTANT, EARTH_RADIUS, EARTH_OMEGA
from . import f90_modules


class QGFieldBase(ABC):
    """Local wave activity and flux analysis in the quasi-geostrophic framework (abstract; instantiate
    ``QGFieldNH18`` or ``QGFieldNHN22``).  Parameters: see falwa.oopinterface.QGFieldBase (oopinterface.py:40-91)."""

    _kind = None  # 0 = NH18, 1 = NHN22

    def __init__(self, xlon, ylat, plev, u_field, v_field, t_field, kmax=49, maxit=100000, dz=1000., npart=None,
                 tol=1.e-5, rjac=0.95, scale_height=SCALE_HEIGHT, cp=CP, dry_gas_constant=DRY_GAS_CONSTANT,
                 omega=EARTH_OMEGA, planet_radius=EARTH_RADIUS, northern_hemisphere_results_only=False,
                 data_on_evenly_spaced_pseudoheight_grid=False, raise_error_for_nonconvergence=True,
                 eq_boundary_index=0):
        self._raise_error_for_nonconvergence = raise_error_for_nonconvergence
        self._nonconvergent_uref = False
        self._data_on_evenly_spaced_pseudoheight_grid = data_on_evenly_spaced_pseudoheight_grid
        # coordinates are widened to float64 whatever their storage type (files hold float32 / int32 coordinates; the
        # reference would carry that precision into its float32 natives, this pipeline is float64 throughout)
        plev = np.asarray(plev, dtype=np.float64)
        if data_on_evenly_spaced_pseudoheight_grid:  # oopinterface.py:110-115
            self.plev = plev
            self.kmax = plev.size
            self._plev_to_height = -scale_height * np.log(plev / P_GROUND)
            self.height = -scale_height * np.log(plev / P_GROUND)
            self.dz = np.diff(self.height)[0]
        else:
            self._check_valid_plev(plev, scale_height, kmax, dz)
            self.kmax = kmax
            self._plev_to_height = -scale_height * np.log(plev / P_GROUND)
            self.height = np.array([i * dz for i in range(self.kmax)])
            self.dz = dz
        u_field = self._convert_masked_data(u_field, "u_field")
        v_field = self._convert_masked_data(v_field, "v_field")
        t_field = self._convert_masked_data(t_field, "t_field")
        ylat = self._convert_masked_data(ylat, "ylat")
        self._input_ylat, self.need_latitude_interpolation, self._ylat, self.equator_idx, self._clat = \
            self._check_and_flip_ylat(np.asarray(ylat, dtype=np.float64))
        self.xlon = np.asarray(xlon, dtype=np.float64)
        self.nlev, self.nlat, self.nlon = plev.size, np.asarray(ylat).size, self.xlon.size
        expected = (self.nlev, self.nlat, self.nlon)
        for field, name in ((u_field, "u_field"), (v_field, "v_field"), (t_field, "theta_field")):
            self._check_dimension_of_fields(field, name, expected)
        if self.need_latitude_interpolation:  # oopinterface.py:148-154 (host regrid of the inputs, rare path)
            u_field = interp1d(self._input_ylat, u_field, axis=1, fill_value="extrapolate")(self._ylat)
            v_field = interp1d(self._input_ylat, v_field, axis=1, fill_value="extrapolate")(self._ylat)
            # the reference regrids theta; theta = T * f(level) and the regrid is linear in T, so regridding T is
            # the same operation up to rounding
            t_field = interp1d(self._input_ylat, t_field, axis=1, fill_value="extrapolate")(self._ylat)
        self._nlat_analysis = self._ylat.size
        self._eq_boundary_index = eq_boundary_index if self._kind == 1 else 0
        self._jd = self._nlat_analysis // 2 + self._nlat_analysis % 2 - self._eq_boundary_index
        self.dphi = np.deg2rad(180. / (self._nlat_analysis - 1))
        self.dlambda = np.deg2rad(360. / self.nlon)
        self.npart = npart if npart is not None else self._nlat_analysis
        self._northern_hemisphere_results_only = northern_hemisphere_results_only
        self.maxit, self.tol, self.rjac = maxit, tol, rjac
        self.scale_height, self.cp, self.dry_gas_constant = scale_height, cp, dry_gas_constant
        self.omega, self.planet_radius = omega, planet_radius
        self._plan = engine.Plan(
            kind=self._kind, xlon=self.xlon, ylat=self._ylat, plev=plev, kmax=self.kmax, dz=self.dz,
            jb=self._eq_boundary_index, npart=self.npart, maxit=maxit, tol=tol, rjac=rjac, scale_height=scale_height,
            cp=cp, dry_gas_constant=dry_gas_constant, omega=omega, planet_radius=planet_radius,
            on_height_grid=data_on_evenly_spaced_pseudoheight_grid)
        self._prefactor = self._plan.prefactor
        self._batch = engine.QGBatch(
            self._plan, np.asarray(u_field)[None], np.asarray(v_field)[None], np.asarray(t_field)[None],
            on_height_grid=data_on_evenly_spaced_pseudoheight_grid,
            northern_hemisphere_results_only=northern_hemisphere_results_only)
        self._reference_states_computed = False
        self._layerwise_fluxes_computed = False
        self._lwa_only = None
        self._cache = {}

    # ---- validation helpers (same messages as the reference) ---------------------------------
    @staticmethod
    def _convert_masked_data(variable, varname):
        if np.ma.is_masked(variable) or isinstance(variable, np.ma.core.MaskedArray):
            warnings.warn(
                '{var} is a masked array of dimension {dim} with {num} masked elements and fill value {fill}. '
                .format(var=varname, dim=variable.shape, num=variable.mask.sum(), fill=variable.fill_value))
            variable = variable.data
        return variable

    def _check_valid_plev(self, plev, scale_height, kmax, dz):
        if np.diff(plev)[0] > 0:
            raise TypeError("plev must be in decending order (i.e. from ground level to aloft)")
        self.plev = plev
        self._plev_to_height = -scale_height * np.log(plev / P_GROUND)
        hmax = -scale_height * np.log(plev[-1] / P_GROUND)
        if hmax < (kmax - 1) * dz:
            raise ValueError('Input kmax = {} but the maximum valid kmax'.format(kmax) +
                             '(constrainted by the vertical grid of your input data) is {}'.format(int(hmax // dz) + 1))

    @staticmethod
    def _check_and_flip_ylat(ylat):
        if np.diff(ylat)[0] < 0:
            raise TypeError("ylat must be in ascending order")
        _input_ylat = ylat
        if (ylat.size % 2 == 0) & (sum(ylat == 0.0) == 0):
            need = True
            _ylat = np.linspace(-90., 90., ylat.size + 1, endpoint=True)
            equator_idx = np.argwhere(_ylat == 0)[0][0] + 1
        elif sum(ylat == 0) == 1:
            need = False
            _ylat = ylat
            equator_idx = np.argwhere(ylat == 0)[0][0] + 1
        else:
            raise TypeError("There are more than 1 grid point with latitude 0.")
        _clat = np.abs(np.cos(np.deg2rad(_ylat)))
        return _input_ylat, need, _ylat, equator_idx, _clat

    @staticmethod
    def _check_dimension_of_fields(field, field_name, expected_dim):
        if np.asarray(field).shape != expected_dim:
            raise TypeError("Incorrect dimension of {}. Expected dimension: {}".format(field_name, expected_dim))

    # ---- output helpers --------------------------------------------------------------------------
    def _interp_back(self, field, interp_from, interp_to, which_axis=1):
        return interp1d(interp_from, field, axis=which_axis, bounds_error=False, fill_value='extrapolate')(interp_to)

    def _return_interp_variables(self, variable, interp_axis):
        if self.need_latitude_interpolation:
            if self.northern_hemisphere_results_only:
                return self._interp_back(variable, self._ylat[-(self._nlat_analysis // 2 + 1):],
                                         self._input_ylat[-(self.nlat // 2):], which_axis=interp_axis)
            return self._interp_back(variable, self._ylat, self._input_ylat, which_axis=interp_axis)
        return variable

    def _host(self, key, tensor):
        """device tensor of batch 1 -> cached NumPy array"""
        if key not in self._cache:
            self._cache[key] = engine.to_host(tensor[0])
        return self._cache[key]

    def _nh_slice(self, a, lat_axis):
        if self.northern_hemisphere_results_only:
            sl = [slice(None)] * a.ndim
            sl[lat_axis] = slice(-self.equator_idx, None)
            return a[tuple(sl)]
        return a

    # ---- the three stages ------------------------------------------------------------------------
    def interpolate_fields(self, return_named_tuple: bool = True) -> Optional[NamedTuple]:
        """Vertical interpolation of u, v, theta onto the uniform pseudo-height grid and QGPV
        (oopinterface.py:467-535).  Returns Interpolated_fields(QGPV, U, V, Theta, Static_stability)."""
        self._cache.clear()
        self._batch.interpolate_fields()
        if return_named_tuple:
            tup = namedtuple('Interpolated_fields', ['QGPV', 'U', 'V', 'Theta', 'Static_stability'])
            return tup(self.qgpv, self.interpolated_u, self.interpolated_v, self.interpolated_theta,
                       self._static_stability_for_tuple())

    def compute_reference_states(self, return_named_tuple: bool = True, northern_hemisphere_results_only=None):
        """Reference states Qref, Uref, PTref (oopinterface.py:550-615)."""
        if northern_hemisphere_results_only:
            warnings.warn(
                "Since v0.7.0, northern_hemisphere_results_only is initialized at the creation of QGField instance. "
                "Please remove this input argument from the method 'compute_reference_states'.")
        if self._batch.qgpv is None:
            raise ValueError("QGField.interpolate_fields has to be called before QGField.compute_reference_states.")
        self._batch.compute_reference_states()
        for k in ("qref", "uref", "ptref", "qref_st"):
            self._cache.pop(k, None)
        self._after_reference_states()
        if not self._nonconvergent_uref:
            self._reference_states_computed = True
        if return_named_tuple:
            Reference_states = namedtuple('Reference_states', ['Qref', 'Uref', 'PTref'])
            return Reference_states(self.qref, self.uref, self.ptref)

    def _after_reference_states(self):
        pass

    def compute_lwa_and_barotropic_fluxes(self, return_named_tuple: bool = True,
                                          northern_hemisphere_results_only=None, ncforce=None, method=0):
        """Barotropic LWA and flux terms (oopinterface.py:796-962).  ``ncforce`` (optional): diabatic term on the
        height grid, shape (kmax, nlat, nlon).  ``method``: 0 automatic, 1 brute-force LWA (extension)."""
        if northern_hemisphere_results_only:
            warnings.warn(
                "Since v0.7.0, northern_hemisphere_results_only is initialized at the creation of QGField instance. "
                "Please remove this input argument from the method 'compute_lwa_and_barotropic_fluxes'.")
        if self._batch.qgpv is None:
            raise ValueError("QGField.interpolate_fields has to be called before QGField.compute_reference_states.")
        if self._nonconvergent_uref:
            raise ValueError(
                """
                Uref was not properly solved. This method cannot be called. 
                Try compute_lwa_only instead if you deem appropriate.
                """)
        if not self._reference_states_computed:
            raise ValueError("Reference states have not been computed yet.")
        if ncforce is not None:
            ncforce = np.asarray(ncforce)
            assert ncforce.shape == (self.kmax, self._nlat_analysis, self.nlon)
            ncforce = ncforce[None]
        self._batch.compute_lwa_and_barotropic_fluxes(ncforce=ncforce, method=method)
        self._lwa_only = None
        for k in list(self._cache):
            if k.startswith("out:") or k == "lwa":
                self._cache.pop(k)
        if return_named_tuple:
            LWA_and_fluxes = namedtuple(
                'LWA_and_fluxes',
                ['adv_flux_f1', 'adv_flux_f2', 'adv_flux_f3', 'convergence_zonal_advective_flux',
                 'divergence_eddy_momentum_flux', 'meridional_heat_flux', 'lwa_baro', 'u_baro', 'lwa'])
            # like the reference, the tuple carries the analysis-grid arrays (no back-interpolation)
            return LWA_and_fluxes(*[self._out2d(n) for n in LWA_and_fluxes._fields[:-1]], self._lwa3d())

    def compute_lwa_only(self) -> None:
        """LWA and its barotropic component from QGPV and Qref only (oopinterface.py:624-690), for the case where the
        Uref solver did not converge.  Goes through the F2PY-compatible entry point compute_lwa_only_nhn22."""
        b = self._batch
        if b.qref_cu is None:
            raise ValueError("Reference states have not been computed yet.")
        eq, nlat = self.equator_idx, self._nlat_analysis
        pv, uu = b.qgpv[0], b.u[0]                       # [k][lat][lon] == Fortran (lon, lat, k)
        qcu = b.qref_cu[0]                               # [k][lat]
        kw = dict(jb=self._eq_boundary_index, is_nhem=True, a=self.planet_radius, om=self.omega, dz=self.dz,
                  h=self.scale_height, rr=self.dry_gas_constant, cp=self.cp, prefac=self.prefactor)
        lwa = np.zeros((self.nlon, nlat, self.kmax), order="F")
        lwa_baro = np.zeros((self.nlon, nlat), order="F")
        u_baro = np.zeros((self.nlon, nlat), order="F")
        ab, ub, a1, a2 = f90_modules.compute_lwa_only_nhn22(pv=pv.contiguous(), uu=uu.contiguous(),
                                                           qref=qcu[:, -eq:].contiguous(), **kw)
        lwa[:, -eq:, :], lwa_baro[:, -eq:], u_baro[:, -eq:] = np.abs(a1 + a2), ab, ub
        # oopinterface.py:666-667 stores |astar1|, |astar2| of the northern hemisphere in the layer-wise storage; the
        # southern assignments (:689-690) go to attributes nothing reads, so those rows stay zero there as well
        p1, p2 = (np.zeros((self.nlon, nlat, self.kmax), order="F") for _ in range(2))
        p1[:, -eq:, :], p2[:, -eq:, :] = np.abs(a1), np.abs(a2)
        if not self.northern_hemisphere_results_only:
            ab, ub, a1, a2 = f90_modules.compute_lwa_only_nhn22(
                pv=(-pv.flip(1)).contiguous(), uu=uu.flip(1).contiguous(),
                qref=(-qcu[:, :eq].flip(1)).contiguous(), **kw)
            lwa[:, :eq, :] = np.abs(a1 + a2)[:, ::-1, :]
            lwa_baro[:, :eq], u_baro[:, :eq] = ab[:, ::-1], ub[:, ::-1]
        self._lwa_only = dict(lwa=np.swapaxes(lwa, 0, 2), lwa_baro=lwa_baro.T, u_baro=u_baro.T,
                              astar1=np.swapaxes(p1, 0, 2), astar2=np.swapaxes(p2, 0, 2))
        # like the reference, whose storages are overwritten (oopinterface.py:676-690): results of an earlier flux
        # stage on this object no longer stand in front of these
        b.lwa = b.out2d = None
        for k in list(self._cache):
            if k == "lwa" or k.startswith("out:"):
                self._cache.pop(k)

    # ---- layer-wise (un-averaged) flux terms: the next step of the path (oopinterface.py:985-1156) ------------
    def compute_ncforce_from_heating_rate(self, heating_rate):
        """q-dot = f e^{z/H} d/dz (e^{-z/H} theta-dot / (d theta/dz)) on the height grid from the diabatic heating rate
        on pressure levels, shape (nlev, nlat, nlon) (oopinterface.py:985-1019, utilities.py:305-344): vertical
        interpolation and the z-derivative in one kernel (falwa_b200_ncforce_from_heating_rate)."""
        if self._batch.profiles is None:
            raise ValueError("QGField.interpolate_fields has to be called first.")
        hr = np.ascontiguousarray(heating_rate, dtype=np.float64)
        return engine.to_host(self._batch.ncforce_from_heating_rate(hr[None])[0])

    def compute_layerwise_lwa_fluxes(self, ncforce=None, method=0):
        """Layer-wise LWA flux terms ua1, ua2, ep1, ep2, ep3, the stretching term and (optionally) ncforce, as 3-D
        fields (oopinterface.py:1063-1156): one batched call (falwa_b200_layerwise_lwa_fluxes) on the device-resident
        fields; ``lwa`` afterwards is astar1 + astar2 like the reference's layer-wise store."""
        b = self._batch
        if b.qref_cu is None or not self._reference_states_computed:
            raise ValueError("Reference states have not been computed yet.")
        nc = None if ncforce is None else np.ascontiguousarray(ncforce, dtype=np.float64)[None]
        out = engine.to_host(b.compute_layerwise_lwa_fluxes(ncforce=nc, method=method)[0])
        self._layerwise = {n: out[i] for n, i in engine.LAYER3D.items()}
        self._layerwise_fluxes_computed = True

    def _layer(self, name):
        if name in ("astar1", "astar2") and not self._layerwise_fluxes_computed and self._lwa_only is not None:
            return self._return_interp_variables(self._nh_slice(self._lwa_only[name], lat_axis=1), interp_axis=1)
        if not self._layerwise_fluxes_computed:
            raise ValueError(f'{name} is not computed yet. Call compute_layerwise_lwa_fluxes() first.')
        return self._return_interp_variables(self._nh_slice(self._layerwise[name], lat_axis=1), interp_axis=1)

    ua1 = property(lambda self: self._layer("ua1"), doc="Layerwise first/linear term of the zonal advective flux")
    ua2 = property(lambda self: self._layer("ua2"), doc="Layerwise second/nonlinear term of the zonal advective flux")
    ep1 = property(lambda self: self._layer("ep1"), doc="Layerwise third term of the zonal advective flux")
    ep2 = property(lambda self: self._layer("ep2"), doc="(u_e v_e cos^2)(j+1), layerwise")
    ep3 = property(lambda self: self._layer("ep3"), doc="(u_e v_e cos^2)(j-1), layerwise")
    stretch_term = property(lambda self: self._layer("stretch_term"), doc="Layerwise stretching (heat-flux) term")
    ncforce = property(lambda self: self._layer("ncforce"), doc="Layerwise non-conservative forcing integral")
    # The reference keeps the two pieces of LWA per layer in its LayerwiseFluxTermsStorage (data_storage.py:199-217;
    # written at oopinterface.py:1085-1086, :1118-1119 and, as magnitudes, by compute_lwa_only at :665-667) without a
    # public accessor; here they can be read like the other layer-wise fields.
    astar1 = property(lambda self: self._layer("astar1"), doc="Layerwise cyclonic part of LWA * cos(lat)")
    astar2 = property(lambda self: self._layer("astar2"), doc="Layerwise anticyclonic part of LWA * cos(lat)")

    @abstractmethod
    def _static_stability_for_tuple(self):
        ...

    # ---- fixed properties --------------------------------------------------------------------------
    @property
    def prefactor(self):
        """Normalization constant for vertical weighted-averaged integration"""
        return self._prefactor

    @property
    def eq_boundary_index(self):
        return self._eq_boundary_index

    @property
    def jd(self):
        return self._jd

    @property
    def ylat(self):
        return self._input_ylat

    @property
    def ylat_ref_states(self):
        if self.northern_hemisphere_results_only:
            if self.need_latitude_interpolation:
                return self._input_ylat[-(self.nlat // 2):]
            return self._input_ylat[-(self.nlat // 2 + 1):]
        return self._input_ylat

    @property
    def ylat_ref_states_analysis(self):
        if self.northern_hemisphere_results_only:
            return self._ylat[-(self._nlat_analysis // 2 + 1):]
        return self._ylat

    @property
    def northern_hemisphere_results_only(self) -> bool:
        return self._northern_hemisphere_results_only

    @property
    def nonconvergent_uref(self) -> bool:
        return self._nonconvergent_uref

    def get_latitude_dim(self):
        return self._input_ylat.size

    # ---- derived quantities ------------------------------------------------------------------------
    def _field3d(self, key, tensor, what):
        if tensor is None:
            raise ValueError(f'{what} is not present in the QGField object.')
        return self._return_interp_variables(self._host(key, tensor), interp_axis=1)

    @property
    def qgpv(self):
        """Quasi-geostrophic potential vorticity on the regular pseudoheight grids, (kmax, nlat, nlon)."""
        return self._field3d("qgpv", self._batch.qgpv, "QGPV field")

    @property
    def interpolated_u(self):
        return self._field3d("u", self._batch.u, "interpolated_u")

    @property
    def interpolated_v(self):
        return self._field3d("v", self._batch.v, "interpolated_v")

    @property
    def interpolated_theta(self):
        return self._field3d("theta", self._batch.theta, "interpolated_theta")

    @property
    def interpolated_avort(self):
        return self._field3d("avort", self._batch.avort, "interpolated_avort")

    @property
    @abstractmethod
    def static_stability(self):
        ...

    def _profile(self, key, tensor, what):
        if tensor is None:
            raise ValueError(f'{what} is not computed yet.')
        a = self._nh_slice(self._host(key, tensor), lat_axis=1)
        return self._return_interp_variables(a, interp_axis=1)

    @property
    def qref(self):
        """Reference state of QGPV (Qref), (kmax, nlat)."""
        return self._profile("qref", self._batch.qref_cu, "qref")

    @property
    def uref(self):
        return self._profile("uref", self._batch.uref, "uref")

    @property
    def ptref(self):
        return self._profile("ptref", self._batch.ptref, "ptref")

    def _out2d(self, name):
        if self._batch.out2d is None:
            if name in ("lwa_baro", "u_baro") and self._lwa_only is not None:
                return self._nh_slice(self._lwa_only[name], lat_axis=0)
            if name in ("lwa_baro", "u_baro", "ncforce_baro"):
                raise ValueError(f'{name} is not computed yet.')
            lat_dim = self.equator_idx if self.northern_hemisphere_results_only else self._nlat_analysis
            return np.zeros((lat_dim, self.nlon))  # the reference's storages are zero-initialised
        return self._nh_slice(self._host("out:" + name, self._batch.out(name)), lat_axis=0)

    def _lwa3d(self):
        if self._batch.lwa is None and self._layerwise_fluxes_computed:
            return self._nh_slice(self._layerwise["lwa"], lat_axis=1)
        if self._batch.lwa is None:
            if self._lwa_only is not None:
                return self._nh_slice(self._lwa_only["lwa"], lat_axis=1)
            raise ValueError('lwa is not computed yet.')
        return self._nh_slice(self._host("lwa", self._batch.lwa), lat_axis=1)

    def _out2d_user(self, name):
        return self._return_interp_variables(self._out2d(name), interp_axis=0)

    adv_flux_f1 = property(lambda self: self._out2d_user("adv_flux_f1"), doc="F1 of eq. 3 of NH18, (nlat, nlon)")
    adv_flux_f2 = property(lambda self: self._out2d_user("adv_flux_f2"), doc="F2 of eq. 3 of NH18")
    adv_flux_f3 = property(lambda self: self._out2d_user("adv_flux_f3"), doc="F3 of eq. 3 of NH18")
    convergence_zonal_advective_flux = property(
        lambda self: self._out2d_user("convergence_zonal_advective_flux"), doc="-div(F1+F2+F3)")
    divergence_eddy_momentum_flux = property(
        lambda self: self._out2d_user("divergence_eddy_momentum_flux"), doc="(II) of eq. 2 of NH18")
    meridional_heat_flux = property(lambda self: self._out2d_user("meridional_heat_flux"), doc="(III) of eq. 2")
    flux_vector_lambda_baro = property(lambda self: self._out2d_user("flux_vector_lambda_baro"))
    flux_vector_phi_baro = property(lambda self: self._out2d_user("flux_vector_phi_baro"))
    lwa_baro = property(lambda self: self._out2d_user("lwa_baro"), doc="barotropic LWA (with cosine weighting)")
    u_baro = property(lambda self: self._out2d_user("u_baro"), doc="barotropic zonal wind")
    ncforce_baro = property(lambda self: self._out2d_user("ncforce_baro"))

    @property
    def lwa(self):
        """Three-dimensional array of local wave activity, (kmax, nlat, nlon)."""
        return self._return_interp_variables(self._lwa3d(), interp_axis=1)


class QGFieldNH18(QGFieldBase):
    """Reference state with the boundary conditions of Nakamura & Huang (2018, Science): red-black SOR
    (oopinterface.py:1546-1687)."""
    _kind = 0

    def _static_stability_for_tuple(self):
        return self.static_stability

    def _after_reference_states(self):
        # oopinterface.py:1617-1631: a hemisphere whose SOR reached maxit is flagged (north checked first)
        nit = self._batch.num_of_iter[0]
        hems = ("Northern",) if self.northern_hemisphere_results_only else ("Northern", "Southern")
        for name, n in zip(hems, nit):
            if n >= self.maxit:
                self._nonconvergent_uref = True
                msg = f"The reference state does not converge for {name} Hemisphere."
                if self._raise_error_for_nonconvergence:
                    raise ValueError(msg)
                warnings.warn(msg)

    @property
    def static_stability(self):
        if self._batch.profiles is None:
            return None
        return self._batch.profiles[0, 2].copy()


class QGField(QGFieldNH18):
    """Alias of ``QGFieldNH18`` kept for backward compatibility (oopinterface.py:1690-1694)."""


class QGFieldNHN22(QGFieldBase):
    """Reference state with the boundary conditions of Neal et al. (2022, GRL): direct inversion
    (oopinterface.py:1697-1904).  ``eq_boundary_index``: grid points between the equator and the domain boundary."""
    _kind = 1

    def __init__(self, xlon, ylat, plev, u_field, v_field, t_field, kmax=49, maxit=100000, dz=1000., npart=None,
                 tol=1.e-5, rjac=0.95, scale_height=SCALE_HEIGHT, cp=CP, dry_gas_constant=DRY_GAS_CONSTANT,
                 omega=EARTH_OMEGA, planet_radius=EARTH_RADIUS, northern_hemisphere_results_only=False,
                 eq_boundary_index=5, data_on_evenly_spaced_pseudoheight_grid=False):
        super().__init__(xlon, ylat, plev, u_field, v_field, t_field, kmax, maxit, dz, npart, tol, rjac, scale_height,
                         cp, dry_gas_constant, omega, planet_radius, northern_hemisphere_results_only,
                         data_on_evenly_spaced_pseudoheight_grid, eq_boundary_index=eq_boundary_index)

    def _static_stability_for_tuple(self):
        p = self._batch.profiles[0]
        return p[2].copy(), p[3].copy()

    @property
    def static_stability(self):
        if self._batch.profiles is None:
            return None
        p = self._batch.profiles[0]
        if self.northern_hemisphere_results_only:
            return p[3].copy()
        return p[2].copy(), p[3].copy()
