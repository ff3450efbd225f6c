from crick_b200.stats import SummaryStats  # noqa: F401
