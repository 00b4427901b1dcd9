"""Importable ``kenlm`` namespace backed by coral_b200's GPU-resident model.

Only what the decode path reaches is provided: ``Model(path)`` with ``order`` / ``path`` /
``in``. Host-side ``BaseScore`` does not exist -- scoring runs inside the
beam-search kernel.
"""
from coral_b200.language_model import KenlmModel as Model  # noqa: F401


class State:  # placeholder: states live on the device
    pass


LanguageModel = Model
