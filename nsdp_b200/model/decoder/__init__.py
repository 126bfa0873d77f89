from nsdp_b200.model.decoder.crosstransformer_decoder import CrossTransformerDecoder
from nsdp_b200.model.decoder.interpolation_decoder import PointInterpDecoder

# same registry as the reference (model/decoder/__init__.py:5-8); 'interp' is its ablation decoder
decoder_dict = {
    "interp": PointInterpDecoder,
    "crossatten": CrossTransformerDecoder,
}
