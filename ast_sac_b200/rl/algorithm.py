"""BatchRLAlgorithm.train for the batched env (ast_sac/core/batch_rl_algorithm.py:47-106,
rl_algorithm.py:57-141): initial exploration, then per epoch evaluation rollouts, exploration rollouts into
the replay buffer and ``num_trains_per_train_loop`` SAC updates.  Wall-clock per phase is kept like the
reference's gtimer stamps ('evaluation sampling', 'exploration sampling', 'training')."""
from __future__ import annotations

import time
from collections import OrderedDict

import torch


class BatchRLAlgorithm:
    def __init__(self, trainer, exploration_data_collector, evaluation_data_collector, replay_buffer, batch_size,
                 max_path_length, num_epochs, num_eval_steps_per_epoch, num_expl_steps_per_train_loop,
                 num_trains_per_train_loop, num_train_loops_per_epoch=1, min_num_steps_before_training=0, log=print,
                 use_cuda_graph=False, logger=None):
        self.trainer = trainer
        self.expl_data_collector, self.eval_data_collector = exploration_data_collector, evaluation_data_collector
        self.replay_buffer = replay_buffer
        self.batch_size, self.max_path_length, self.num_epochs = batch_size, max_path_length, num_epochs
        self.num_eval_steps_per_epoch = num_eval_steps_per_epoch
        self.num_expl_steps_per_train_loop = num_expl_steps_per_train_loop
        self.num_trains_per_train_loop = num_trains_per_train_loop
        self.num_train_loops_per_epoch = num_train_loops_per_epoch
        self.min_num_steps_before_training = min_num_steps_before_training
        self.log = log
        self.use_cuda_graph = use_cuda_graph
        # optional rl/logging.Logger: one progress.csv row and one snapshot per epoch, in the reference's format
        self.logger = logger
        self._t_start = time.perf_counter()
        self._t_logging = 0.0         # duration of the previous epoch's logging (gtimer stamps it in the next epoch)
        self.history = []

    def _get_snapshot(self):
        """rl_algorithm.py:76-86."""
        snapshot = {}
        for k, v in self.trainer.get_snapshot().items():
            snapshot['trainer/' + k] = v
        for k, v in self.expl_data_collector.get_snapshot().items():
            snapshot['exploration/' + k] = v
        for k, v in self.eval_data_collector.get_snapshot().items():
            snapshot['evaluation/' + k] = v
        for k, v in self.replay_buffer.get_snapshot().items():
            snapshot['replay_buffer/' + k] = v
        return snapshot

    def _log_stats(self, epoch, times):
        """rl_algorithm.py:88-141: the same record_dict calls, prefixes and order; `times` holds this epoch's phase
        durations under gtimer's stamp names."""
        lg = self.logger
        lg.log("Epoch {} finished".format(epoch), with_timestamp=True)
        lg.record_dict({"epoch": epoch})
        lg.record_dict(self.replay_buffer.get_diagnostics(), prefix='replay_buffer/')
        lg.record_dict(self.trainer.get_diagnostics(), prefix='trainer/')
        lg.record_dict(self.expl_data_collector.get_diagnostics(), prefix='expl/')
        lg.record_dict(self.expl_data_collector.get_generic_path_information(), prefix='expl/')
        lg.record_dict(self.eval_data_collector.get_diagnostics(), prefix='eval/')
        lg.record_dict(self.eval_data_collector.get_generic_path_information(), prefix='eval/')
        t = OrderedDict()
        epoch_time = 0.0
        for key in sorted(times):
            epoch_time += times[key]
            t['time/{} (s)'.format(key)] = times[key]
        t['time/epoch (s)'] = epoch_time
        t['time/total (s)'] = time.perf_counter() - self._t_start
        lg.record_dict(t)
        lg.record_tabular('Epoch', epoch)
        lg.dump_tabular(with_prefix=False, with_timestamp=False)

    def _sync(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def train(self):
        times = OrderedDict()
        if self.min_num_steps_before_training > 0:
            t0 = time.perf_counter()
            self.expl_data_collector.collect_new_steps(self.max_path_length, self.min_num_steps_before_training, False)
            self.expl_data_collector.end_epoch(-1)
            self._sync()
            times['initial exploration (s)'] = time.perf_counter() - t0
        for epoch in range(self.num_epochs):
            t0 = time.perf_counter()
            self.eval_data_collector.collect_new_steps(self.max_path_length, self.num_eval_steps_per_epoch, True)
            self._sync()
            t1 = time.perf_counter()
            t_expl = t_train = 0.0
            for _ in range(self.num_train_loops_per_epoch):
                ta = time.perf_counter()
                self.expl_data_collector.collect_new_steps(self.max_path_length, self.num_expl_steps_per_train_loop, False)
                self._sync()
                tb = time.perf_counter()
                if self.use_cuda_graph and self.trainer._graph is None:
                    self.trainer.capture(self.replay_buffer, self.batch_size)
                for _ in range(self.num_trains_per_train_loop):
                    if self.use_cuda_graph:
                        self.trainer.train_graphed()
                    else:
                        self.trainer.train_from_torch(self.replay_buffer.random_batch(self.batch_size))
                self._sync()
                tc = time.perf_counter()
                t_expl += tb - ta
                t_train += tc - tb
            stats = OrderedDict(epoch=epoch)
            stats['time/evaluation sampling (s)'] = t1 - t0
            stats['time/exploration sampling (s)'] = t_expl
            stats['time/training (s)'] = t_train
            for k, v in self.eval_data_collector.get_diagnostics().items():
                stats['evaluation/' + k] = v
            for k, v in self.expl_data_collector.get_diagnostics().items():
                stats['exploration/' + k] = v
            for k, v in self.trainer.get_diagnostics().items():
                stats['trainer/' + k] = v
            stats['replay_buffer/size'] = self.replay_buffer.num_steps_can_sample()
            stats.update(times)
            times = OrderedDict()
            self.history.append(stats)
            self.log(stats)
            if self.logger is not None:
                ts = time.perf_counter()
                self.logger.save_itr_params(epoch, self._get_snapshot())
                t_save = time.perf_counter() - ts
                # gtimer stamps of batch_rl_algorithm.py:47-106 ('data storing' is part of the rollout here: the
                # transitions go straight into the GPU replay buffer; 'training' is the loop around the updates)
                self._log_stats(epoch, {'evaluation sampling': t1 - t0, 'exploration sampling': t_expl,
                                        'data storing': 0.0, 'sac training': t_train, 'training': 0.0,
                                        'saving': t_save, 'logging': self._t_logging})
                self._t_logging = time.perf_counter() - ts - t_save
            self.eval_data_collector.end_epoch(epoch)
            self.expl_data_collector.end_epoch(epoch)
        return self.history
