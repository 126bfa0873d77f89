"""Drop-in for the reference's pybind11 module `pointnet2_ops._ext`
(pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19): same nine function names, argument order and
return types, implemented by the sm_100a kernels in libnsdp_b200.so through nsdp_b200.ops."""
from nsdp_b200.ops import (ball_query, furthest_point_sampling, gather_points, gather_points_grad, group_points,  # noqa: F401
                           group_points_grad, three_interpolate, three_interpolate_grad, three_nn)
