"""DataInput batch iterator.

The reference imports `DataHandle.get_input_data.DataInput` (train_process.py:12, used at :240, :326)
but the module is missing from its tree; this supplies the contract its call sites rely on:
`for step_i, batch in DataInput(data, batch_size)` yields consecutive slices including a short
last batch.
"""


class DataInput:
    def __init__(self, data, batch_size):
        self.data = data
        self.batch_size = int(batch_size)
        self.epoch_size = (len(data) + self.batch_size - 1) // self.batch_size
        self.i = 0

    def __iter__(self):
        self.i = 0
        return self

    def __len__(self):
        return self.epoch_size

    def __next__(self):
        if self.i >= self.epoch_size:
            raise StopIteration
        ts = self.data[self.i * self.batch_size:(self.i + 1) * self.batch_size]
        self.i += 1
        return self.i, ts
