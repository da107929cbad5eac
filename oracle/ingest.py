"""TEST INFRASTRUCTURE (CPU restatement; never imported by the product path).

What happens to archived fields before the reference's QGField sees them:

* xarray's CF decoding of packed variables (``mask_and_scale``): ``value = float64(raw) * scale_factor + add_offset``
  with the product rounded before the sum, ``raw == _FillValue`` -> NaN.  The reference's tests decode the packed
  ERA-Interim snapshot this way (tests/test_output_results.py:26-45 via ``xr.open_dataset``).
* ``QGDataset`` reorders every field so that latitude ascends and pressure descends
  (src/falwa/xarrayinterface.py:384-405: ``np.flip(field, axis=flip)``).

Parity pinned by: ``oracle.testdata.load_era_snapshot`` (the decoded snapshot every golden fixture of the 3-D path was
generated from) must be reproduced bit for bit from the raw int16 values of the same file.
"""
import numpy as np


def decode_field(raw, scale_factor=1.0, add_offset=0.0, fill_value=None, flip_lat=False, flip_lev=False):
    """raw [..., level, lat, lon] (int16 / float32 / float64) -> float64, decoded and reordered."""
    raw = np.asarray(raw)
    x = raw.astype(np.float64)
    if scale_factor != 1.0 or add_offset != 0.0 or raw.dtype != np.float64:
        x = x * np.float64(scale_factor)
        x = x + np.float64(add_offset)
    if fill_value is not None:
        x[raw.astype(np.float64) == np.float64(fill_value)] = np.nan
    if flip_lat:
        x = np.flip(x, axis=-2)
    if flip_lev:
        x = np.flip(x, axis=-3)
    return np.ascontiguousarray(x)
