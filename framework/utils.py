from gnndelete_b200.kg import negative_sampling_kg  # noqa: F401  (reference: framework/utils.py:46-58)
