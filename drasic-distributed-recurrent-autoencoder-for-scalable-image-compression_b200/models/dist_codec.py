"""Distributed codec: one encoder per data source, one jointly trained decoder
(reference models/dist_codec.py)."""
import config
from .codec_base import CodecModel, decoder_spec, encoder_spec

__all__ = ['dist_codec']


class Model(CodecModel):
    separate = False


def dist_codec():
    """zero-argument factory reading config.PARAM (dist_codec.py:68-106); also writes PARAM['model']"""
    M = config.PARAM['num_node']
    config.PARAM['model'] = {'encoder': [encoder_spec(head_bias=True) for _ in range(M)],
                             'quantizer': {'cell': 'QuantizationCell'},
                             'decoder': decoder_spec()}
    return Model()
