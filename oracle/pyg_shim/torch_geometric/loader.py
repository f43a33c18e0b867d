from .data import Batch


class DataLoader:
    """batch_size / shuffle=False only (enough for golden generation)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        assert not shuffle
        self.dataset, self.batch_size = dataset, batch_size

    def __iter__(self):
        items = [self.dataset[i] for i in range(len(self.dataset))]
        for s in range(0, len(items), self.batch_size):
            yield Batch.from_data_list(items[s:s + self.batch_size])

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size
