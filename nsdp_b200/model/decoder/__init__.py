from nsdp_b200.model.decoder.crosstransformer_decoder import CrossTransformerDecoder

# 'interp' (PointInterpDecoder, model/decoder/interpolation_decoder.py) is an ablation no shipped config
# selects; out of the hot-path scope.
decoder_dict = {
    "crossatten": CrossTransformerDecoder,
}
