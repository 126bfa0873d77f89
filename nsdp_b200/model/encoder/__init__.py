from nsdp_b200.model.encoder.pointnetplusplus import PointNetPlusPlusEncoder
from nsdp_b200.model.encoder.pointransformer import PointTransformerEncoder

# same registry as the reference (model/encoder/__init__.py:4-7); 'pointnet++' is its ablation encoder
encoder_dict = {
    "pointnet++": PointNetPlusPlusEncoder,
    "pointransformer": PointTransformerEncoder,
}
