from coral_b200.alphabet import (  # noqa: F401
    BLANK_TOKEN_PTN, BPE_TOKEN, UNK_BPE_TOKEN, UNK_TOKEN, UNK_TOKEN_PTN, Alphabet,
)
