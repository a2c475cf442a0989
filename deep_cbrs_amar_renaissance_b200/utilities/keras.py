"""Parameter counting and the training log callback (counterparts of /root/reference/src/utilities/keras.py:10-90).
The callback is the reference's only timing instrumentation: mean wall time of training batches 500-520 and
total training time; here they are kept on the object (and written to metrics.json by the driver) instead of MLflow."""
import time

import numpy as np


def _device_sync():
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def get_total_parameters(model):
    """(trainable, non-trainable) parameter counts (keras.py:10-22)"""
    return (int(sum(w.numel() for w in model.trainable_weights)), int(sum(w.numel() for w in model.non_trainable_weights)))


class LogCallback:
    def __init__(self, log, frequency, batch_slice=(500, 520)):
        self.log, self.log_frequency, self.batch_slice = log, frequency, batch_slice
        self.batch_times, self.epoch_logs = [], []
        self.train_start = self.training_time = self._t0 = None
        self.trace = False
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        self.train_start = time.perf_counter()
        self.log.info("Starting training - got log keys: {}".format(list((logs or {}).keys())))

    def on_train_end(self, logs=None):
        self.training_time = time.perf_counter() - self.train_start
        self.log.info("End training")

    def on_epoch_begin(self, epoch, logs=None):
        self.log.info('Epoch: {} '.format(epoch))

    def on_epoch_end(self, epoch, logs=None):
        logs = dict(logs or {})
        self.epoch_logs.append(logs)
        self.log.info("End epoch {} of training - {}".format(epoch, "".join("{}: {},\t".format(k, v) for k, v in logs.items())))

    def on_train_batch_begin(self, batch, logs=None):
        # keras.py:67-76: tracing starts at the first batch whose index falls in the slice and - because the flag that
        # would close it is only looked at while tracing is off - never stops: every later batch of the run is timed
        if self.trace or self.batch_slice[0] <= batch <= self.batch_slice[1]:
            self.trace = True
            _device_sync()
            self._t0 = time.perf_counter()

    def on_train_batch_end(self, batch, logs=None):
        if self.trace and self._t0 is not None:
            _device_sync()  # the step is enqueued asynchronously; without this the host time says nothing
            self.batch_times.append(time.perf_counter() - self._t0)

    def get_batch_time(self):
        return float(np.mean(self.batch_times)) if self.batch_times else float("nan")
