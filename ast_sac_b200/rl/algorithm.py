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
                 use_cuda_graph=False):
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
        self.history = []

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
            self.eval_data_collector.end_epoch(epoch)
            self.expl_data_collector.end_epoch(epoch)
        return self.history
