"""B200-native x-vector embedding extractor (hot path of BUTSpeechFIT/x-vector-kaldi-tf).

Import as ``xvector_b200`` (see ``xvector_b200/__init__.py``); modules mirror the reference's
``local/tf`` file names: ``models``, ``kaldi_io``, ``extract_embedding``, ``ze_utils``.
"""
