from coral_b200.decoder import *  # noqa: F401,F403
from coral_b200.decoder import BeamSearchDecoderCTC, build_ctcdecoder  # noqa: F401
