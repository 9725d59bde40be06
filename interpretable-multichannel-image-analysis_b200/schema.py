"""Column schema of the feature table.

The order is the reference notebook's dict insertion order (NB:330-331, NB:334): the 17 basic
features for Ch1..ChC (NB:241-262; note ``kurtosis`` precedes ``skew`` and the entropy key has
no ``_intensity``), then the 6 GLCM features for Ch1..ChC (NB:301-306).  The channel suffix is
positional and 1-based (``"_Ch" + str(ch+1)``, NB:241).  Extension blocks follow.
"""

BASIC_NAMES = (
    ["min_intensity"]
    + ["percentile%d0_intensity" % k for k in range(1, 10)]
    + ["max_intensity", "total_intensity", "mean_intensity", "std_intensity",
       "kurtosis_intensity", "skew_intensity", "shannon_entropy"]
)
GLCM_PROPS = ["contrast", "dissimilarity", "homogeneity", "ASM", "energy", "correlation"]
ANGLE_TAGS = ["", "_a45", "_a90", "_a135"]
SHAPE_NAMES = [
    "area", "perimeter", "bbox_area", "extent", "centroid_row", "centroid_col",
    "major_axis_length", "minor_axis_length", "eccentricity", "circularity",
]
MOMENT_NAMES = [
    "weighted_centroid_row", "weighted_centroid_col",
    "nu20", "nu11", "nu02", "nu30", "nu21", "nu12", "nu03",
]
NOTEBOOK_PERCENTILES = tuple(k / 10.0 for k in range(1, 10))   # NB:242-250 (0.1 .. 0.9 percent)


def block_widths(glcm=True, n_angles=1, shape=False, moments=False, basic=True):
    """Per-channel widths of the (basic, glcm, shape, moments) blocks."""
    return (len(BASIC_NAMES) if basic else 0,
            len(GLCM_PROPS) * n_angles if glcm else 0,
            len(SHAPE_NAMES) if shape else 0,
            len(MOMENT_NAMES) if moments else 0)


def feature_columns(n_channels, glcm=True, n_angles=1, shape=False, moments=False, basic=True,
                    channel_ids=None):
    """Column names for ``n_channels`` slots.  ``channel_ids`` (1-based ints) overrides the
    positional suffix, e.g. to keep original channel numbers in a leave-one-out table."""
    ids = list(range(1, n_channels + 1)) if channel_ids is None else list(channel_ids)
    assert len(ids) == n_channels
    cols = []
    if basic:
        for k in ids:
            cols += ["%s_Ch%d" % (nm, k) for nm in BASIC_NAMES]
    if glcm:
        for k in ids:
            for tag in ANGLE_TAGS[:n_angles]:
                cols += ["%s%s_Ch%d" % (p, tag, k) for p in GLCM_PROPS]
    if shape:
        for k in ids:
            cols += ["%s_Ch%d" % (nm, k) for nm in SHAPE_NAMES]
    if moments:
        for k in ids:
            cols += ["%s_Ch%d" % (nm, k) for nm in MOMENT_NAMES]
    return cols


def channel_column_index(n_channels, slot, glcm=True, n_angles=1, shape=False, moments=False,
                         basic=True):
    """Indices of every column that belongs to channel slot ``slot`` (0-based)."""
    wb, wg, ws, wm = block_widths(glcm, n_angles, shape, moments, basic)
    idx, base = [], 0
    for wdt in (wb, wg, ws, wm):
        idx += list(range(base + slot * wdt, base + (slot + 1) * wdt))
        base += wdt * n_channels
    return idx
