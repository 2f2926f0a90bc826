"""Mirror of the reference's ``asr`` package for the path this repository replaces: ``asr.loss``, ``asr.error`` and the
label half of ``asr.data.processing``."""
