"""Drop-in `pointnet2_ops` package (reference: pointnet2_ops_lib/pointnet2_ops/__init__.py:1-3) on the nsdp_b200 kernels."""
from nsdp_b200.pointnet2_ops import _ext, pointnet2_utils  # noqa: F401
from nsdp_b200.pointnet2_ops import pointnet2_modules  # noqa: F401,E402
from nsdp_b200.pointnet2_ops._version import __version__  # noqa: F401
