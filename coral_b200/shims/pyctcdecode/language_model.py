from coral_b200.language_model import LanguageModel, load_unigram_set_from_arpa  # noqa: F401
