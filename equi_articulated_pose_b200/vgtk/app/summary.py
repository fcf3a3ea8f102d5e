class Summary():
    """Exponential-moving-average metric book (reference: vgtk/vgtk/app/summary.py:3-27)."""

    def __init__(self, decay=0.9):
        self.decay = decay
        self.values = {}

    def register(self, keys):
        for k in keys:
            self.values.setdefault(k, None)

    def update(self, updates):
        for k, v in updates.items():
            old = self.values.get(k)
            self.values[k] = v if old is None else self.decay * old + (1 - self.decay) * v

    def get(self):
        return "  ".join(f"{k}: {v:.4f}" for k, v in self.values.items() if v is not None)

    def __getitem__(self, k):
        return self.values[k]
