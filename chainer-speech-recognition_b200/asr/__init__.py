"""Mirror of the reference's ``asr`` package for the one path this repository replaces: ``asr.loss``."""
