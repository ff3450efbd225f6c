from crick_b200.space_saving import SpaceSaving, TopKResult  # noqa: F401
