"""Host-side mirrors of the reference's simulator classes.

They keep the reference's class names, constructor signatures and attribute names
(ShipConfiguration ... ShipAssets) so existing set-up code only changes its imports, but they hold
*parameters and initial values only*: all dynamics run in the CUDA kernels of csrc/.  There is no
CPU stepping path in this package.
"""
