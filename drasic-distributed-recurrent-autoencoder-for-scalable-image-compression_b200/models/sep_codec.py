"""Separate codecs: an independent encoder/decoder pair per data source, residual and outputs
detached (reference models/sep_codec.py)."""
import config
from .codec_base import CodecModel, decoder_spec, encoder_spec

__all__ = ['sep_codec']


class Model(CodecModel):
    separate = True


def sep_codec():
    """sep_codec.py:69-107"""
    M = config.PARAM['num_node']
    config.PARAM['model'] = {'encoder': [encoder_spec(head_bias=False) for _ in range(M)],
                             'quantizer': {'cell': 'QuantizationCell'},
                             'decoder': [decoder_spec() for _ in range(M)]}
    return Model()
