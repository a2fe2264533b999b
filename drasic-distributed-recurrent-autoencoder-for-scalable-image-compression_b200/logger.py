"""Running-mean logger with the interface the reference's loops use (logger.py:7-80: safe / reset / append /
write / tracker).  TensorBoard output is optional: it is written only when torch.utils.tensorboard imports."""
from collections import defaultdict
from numbers import Number


class Logger(object):
    def __init__(self, log_path):
        self.log_path = log_path
        self.writer = None
        self.tracker, self.counter, self.mean = defaultdict(int), defaultdict(int), defaultdict(int)
        self.history, self.iterator = defaultdict(list), defaultdict(int)

    def safe(self, write):
        if write:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(self.log_path)
            except Exception:                    # tensorboard not installed: keep logging to stdout only
                self.writer = None
        elif self.writer is not None:
            self.writer.close()
            self.writer = None

    def reset(self):
        for name, value in self.mean.items():
            self.history[name].append(value)
        self.tracker, self.counter, self.mean = defaultdict(int), defaultdict(int), defaultdict(int)

    def append(self, result, tag, n=1, mean=True):
        for k, v in result.items():
            name = '{}/{}'.format(tag, k)
            self.tracker[name] = v
            self.counter[name] += n
            if not mean:
                continue
            seen = self.counter[name]
            if isinstance(v, Number):
                self.mean[name] = ((seen - n) * self.mean[name] + n * v) / seen
            elif isinstance(v, list):
                prev = self.mean[name] if seen != n else [0] * len(v)
                self.mean[name] = [((seen - n) * p + n * x) / seen for p, x in zip(prev, v)]
            else:
                raise ValueError('Not valid data type')

    def write(self, tag, metric_names):
        parts = []
        for k in metric_names:
            name = '{}/{}'.format(tag, k)
            m = self.mean[name]
            if isinstance(m, list):
                m = sum(m) / len(m)
            elif not isinstance(m, Number):
                raise ValueError('Not valid data type')
            parts.append('{}: {:.4f}'.format(k, m))
            if self.writer is not None:
                self.iterator[name] += 1
                self.writer.add_scalar(name, m, self.iterator[name])
        info = list(self.tracker['{}/info'.format(tag)] or [])
        info[2:2] = parts
        line = '  '.join(info)
        print(line)
        if self.writer is not None:
            name = '{}/info'.format(tag)
            self.iterator[name] += 1
            self.writer.add_text(name, line, self.iterator[name])

    def flush(self):
        if self.writer is not None:
            self.writer.flush()

    def __getstate__(self):                      # checkpoints pickle the logger (train_model.py:92)
        d = dict(self.__dict__)
        d['writer'] = None
        return d
