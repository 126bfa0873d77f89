from nsdp_b200.model.encoder.pointransformer import PointTransformerEncoder

# The reference also registers 'pointnet++' (model/encoder/__init__.py:4-7), an ablation no shipped config
# selects; it is out of the hot-path scope and deliberately absent here.
encoder_dict = {
    "pointransformer": PointTransformerEncoder,
}
