"""Mirror of the reference's ``asr`` package for the path this repository replaces: ``asr.loss`` and ``asr.error``."""
