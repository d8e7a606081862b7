"""Stand-in so the reference orchestrator imports; only instantiated for formula strings."""


class FormulaicContrasts:  # pragma: no cover - never instantiated (design passed as DataFrame)
    def __init__(self, *a, **k):
        raise NotImplementedError("formula designs are not available in the shimmed oracle")
