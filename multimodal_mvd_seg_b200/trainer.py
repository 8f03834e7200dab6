"""nnUNetTrainer-style trainers for the built hot path.

``nnUNetTrainer`` keeps the hook names, signatures and return conventions of the reference base trainer
(nnunetv2/training/nnUNetTrainer/nnUNetTrainer.py:68-1004) for the path north_star names: ``initialize``,
``build_network_architecture`` (static), ``_build_loss``, ``_get_deep_supervision_scales``, ``configure_optimizers``,
``set_deep_supervision_enabled``, ``train_step(batch) -> {'loss': np.ndarray}``,
``validation_step(batch) -> {'loss','tp_hard','fp_hard','fn_hard'}``, ``save_checkpoint`` / ``load_checkpoint`` keys.
``MVDTrainer`` (the reference's ``ContrastiveTrainer``, MVDTrainer.py:76-985) adds the second modality network, the
mutual-distillation KL and the topological term with the canonical decisions of SURVEY.md section 8c.

The epoch loop around the step is kept as well (``run_training`` and its ``on_*`` hooks, MVDTrainer.py:808-1127 /
:1323-1345: loss / pseudo-Dice aggregation, EMA, checkpoint cadence), driven by loaders the caller supplies.  Dataset
unpacking, batchgenerators augmentation, file logging, plotting and the sliding-window validation export are out of
scope; ``plans`` / ``dataset_json`` are read through the two small accessors below, which
expose the same property names as the reference's ConfigurationManager / LabelManager
(utilities/plans_handling/plans_handler.py:32-291, utilities/label_handling/label_handling.py:21-234).
"""
import os
import time
from typing import List, Optional, Tuple, Union

import numpy as np
import torch
import torch.distributed as dist
from torch import nn

from . import ops
from .ddp import GradArena, broadcast_parameters, split_batch_for_rank
from .ds_targets import downsample_seg_for_ds
from .losses import (DC_and_CE_loss, DeepSupervisionWrapper, MemoryEfficientSoftDiceLoss, distill_kl, soft_cldice,
                     softmax_channel)
from .network import get_network_from_plans
from .optim import PolyLRScheduler, SGDNesterovClip


class ConfigurationManager(object):
    """property names of plans_handler.py:32-176 (the subset the hot path reads)."""

    def __init__(self, configuration_dict: dict):
        self.configuration = configuration_dict

    def __getattr__(self, name):
        try:
            return self.__dict__['configuration'][name]
        except KeyError:
            raise AttributeError(name)

    @property
    def batch_dice(self) -> bool:
        return self.configuration.get('batch_dice', False)

    @property
    def UNet_class_name(self) -> str:
        return self.configuration.get('UNet_class_name', 'PlainConvUNet')

    @property
    def previous_stage_name(self):
        return self.configuration.get('previous_stage')


class LabelManager(object):
    """label_handling.py:21-234, classes-only subset (no regions, no ignore label on this path)."""

    def __init__(self, label_dict: dict):
        self.label_dict = label_dict
        vals = []
        for v in label_dict.values():
            if isinstance(v, (tuple, list)):
                raise NotImplementedError('region-based training is outside the built hot path')
            vals.append(int(v))
        if 'ignore' in label_dict:
            raise NotImplementedError('the ignore label is outside the built hot path')
        self.all_labels = sorted(vals)
        self.has_regions = False
        self.ignore_label = None
        self.has_ignore_label = False

    @property
    def foreground_labels(self):
        return [i for i in self.all_labels if i != 0]

    @property
    def num_segmentation_heads(self) -> int:
        return len(self.all_labels)


class PlansManager(object):
    def __init__(self, plans: dict):
        self.plans = plans

    def get_configuration(self, name: str) -> ConfigurationManager:
        cfg = dict(self.plans['configurations'][name])
        while 'inherits_from' in cfg:   # plans_handler.py:197-228
            parent = dict(self.plans['configurations'][cfg.pop('inherits_from')])
            parent.update(cfg)
            cfg = parent
        return ConfigurationManager(cfg)

    def get_label_manager(self, dataset_json: dict) -> LabelManager:
        return LabelManager(dataset_json['labels'])


def determine_num_input_channels(plans_manager, configuration_manager, dataset_json) -> int:
    """label_handling.py:283-300 (no cascade on this path)."""
    return len(dataset_json['modality']) if 'modality' in dataset_json else len(dataset_json['channel_names'])


def make_plans(patch_size, batch_size: int = 2, n_modalities: int = 2, n_classes: int = 4, batch_dice: bool = False,
               base_features: int = 32, max_features: int = 320, min_edge: int = 4) -> Tuple[dict, dict]:
    """synthetic plans.json / dataset.json for a 3d_fullres configuration with the reference's topology rule
    (network_topology.py:30-105 for isotropic spacing; default_experiment_planner.py:50-66 constants)."""
    cur = list(patch_size)
    pool = [[1, 1, 1]]
    while True:
        valid = [i for i in range(3) if cur[i] >= 2 * min_edge]
        if len(valid) < 1:
            break
        if len(valid) == 1 and cur[valid[0]] < 3 * min_edge:
            break
        pool.append([2 if i in valid else 1 for i in range(3)])
        cur = [int(np.ceil(c / 2)) if i in valid else c for i, c in enumerate(cur)]
    n = len(pool)
    cfg = {'patch_size': list(patch_size), 'batch_size': batch_size, 'UNet_class_name': 'PlainConvUNet',
           'UNet_base_num_features': base_features, 'unet_max_num_features': max_features,
           'n_conv_per_stage_encoder': [2] * n, 'n_conv_per_stage_decoder': [2] * (n - 1),
           'pool_op_kernel_sizes': pool, 'conv_kernel_sizes': [[3, 3, 3]] * n, 'batch_dice': batch_dice,
           'spacing': [1.0, 1.0, 1.0]}
    plans = {'plans_name': 'syntheticPlans', 'dataset_name': 'Dataset000_Synthetic',
             'configurations': {'3d_fullres': cfg}}
    dataset_json = {'channel_names': {str(i): f'mod{i}' for i in range(n_modalities)},
                    'labels': {'background': 0, **{f'class{i}': i for i in range(1, n_classes)}}}
    return plans, dataset_json


def collate_outputs(outputs: List[dict]) -> dict:
    """merge the per-step result dicts of one epoch (contract of utilities/collate_outputs.py:6-24): python scalars
    become a list, numpy arrays are stacked along a new leading axis, lists are concatenated."""
    def merge(values):
        first = values[0]
        if np.isscalar(first):
            return list(values)
        if isinstance(first, np.ndarray):
            return np.vstack([v[None] for v in values])
        if isinstance(first, list):
            merged = []
            for v in values:
                merged.extend(v)
            return merged
        raise ValueError(f'Cannot collate input of type {type(first)}. Modify collate_outputs to add this functionality')
    return {key: merge([o[key] for o in outputs]) for key in outputs[0]}


class nnUNetLogger(object):
    """per-epoch series under the key names of training/logging/nnunet_logger.py:18-27 (they are part of the
    checkpoint, MVDTrainer.py:1140); no plotting.  Writing 'mean_fg_dice' also extends 'ema_fg_dice' with
    0.9 * previous + 0.1 * value (:49-52)."""
    KEYS = ('mean_fg_dice', 'ema_fg_dice', 'dice_per_class_or_region', 'train_losses', 'val_losses', 'lrs',
            'epoch_start_timestamps', 'epoch_end_timestamps')

    def __init__(self, verbose: bool = False):
        self.my_fantastic_logging = {k: [] for k in self.KEYS}
        self.verbose = verbose

    def log(self, key, value, epoch: int):
        series = self.my_fantastic_logging.get(key)
        assert isinstance(series, list), f'unknown logging key {key}'
        if self.verbose:
            print(f'logging {key}: {value} for epoch {epoch}')
        assert len(series) in (epoch, epoch + 1), 'an epoch was skipped: one entry per epoch and key'
        if len(series) == epoch:
            series.append(value)
        else:
            series[epoch] = value
        if key == 'mean_fg_dice':
            ema = self.my_fantastic_logging['ema_fg_dice']
            self.log('ema_fg_dice', value if not ema else 0.9 * ema[epoch - 1] + 0.1 * value, epoch)

    def get_checkpoint(self):
        return self.my_fantastic_logging

    def load_checkpoint(self, checkpoint: dict):
        self.my_fantastic_logging = checkpoint


def _mean_over_ranks(per_step_values, is_ddp: bool) -> float:
    """mean of a per-step quantity over all steps of all ranks (MVDTrainer.py:990-993, 1083-1088)."""
    if not is_ddp:
        return float(np.mean(per_step_values))
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, per_step_values)
    return float(np.vstack(gathered).mean())


def _sum_over_ranks(counts: np.ndarray, is_ddp: bool) -> np.ndarray:
    if not is_ddp:
        return counts
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, counts)
    return np.sum(np.stack(gathered), axis=0)


class nnUNetTrainer(object):
    def __init__(self, plans: dict, configuration: str, fold: int, dataset_json: dict, unpack_dataset: bool = True,
                 device: torch.device = torch.device('cuda'), specified_cfg: str = ''):
        self.is_ddp = dist.is_available() and dist.is_initialized()
        self.local_rank = 0 if not self.is_ddp else dist.get_rank()
        self.device = device
        if self.device.type != 'cuda':
            raise RuntimeError('this trainer drives hand-written sm_100a kernels: device must be CUDA '
                               '(the reference\'s CPU path is timed by bench.py --impl reference)')
        self.plans_manager = PlansManager(plans)
        self.configuration_manager = self.plans_manager.get_configuration(configuration)
        self.configuration_name = configuration
        self.dataset_json = dataset_json
        self.fold = fold
        self.label_manager = self.plans_manager.get_label_manager(dataset_json)
        # hyper-parameters: MVDTrainer.py:161-166
        self.initial_lr = 1e-2
        self.weight_decay = 3e-5
        self.oversample_foreground_percent = 0.33
        self.num_iterations_per_epoch = 250
        self.num_val_iterations_per_epoch = 50
        self.num_epochs = 1000
        self.current_epoch = 0
        self.enable_deep_supervision = True
        self.num_input_channels = None
        self.network = None
        self.optimizer = self.lr_scheduler = None
        self.grad_scaler = None  # bf16 path: the reference's no-scaler branch (nnUNetTrainer.py:921-924)
        self.loss = None
        self.was_initialized = False
        self.use_cuda_graph = False
        self.split_graph = False   # capture forward and (loss + backward + optimiser) as two graphs: the H2D copy of
                                   # the targets then overlaps the forward pass (see train_step_async)
        self.graph_warmup_steps = 2
        # epoch loop (run_training, MVDTrainer.py:1323-1345): the loaders are set by the caller -- batch production is
        # outside this package -- and checkpoints are only written when an output folder is given
        self.logger = nnUNetLogger()
        self.dataloader_train = self.dataloader_val = None
        self.output_folder: Optional[str] = None
        self.save_every = 50
        self._best_ema = None
        self._arenas: List[GradArena] = []
        self._set_batch_size_and_oversample()

    # ------------------------------------------------------------------------------------------------------------
    def _set_batch_size_and_oversample(self):
        if not self.is_ddp:
            self.batch_size = self.configuration_manager.batch_size
        else:
            self.batch_size, self.oversample_foreground_percent = split_batch_for_rank(
                self.configuration_manager.batch_size, dist.get_world_size(), dist.get_rank(),
                self.oversample_foreground_percent)

    def initialize(self):
        if self.was_initialized:
            raise RuntimeError('You have called self.initialize even though the trainer was already initialized.')
        self.num_input_channels = determine_num_input_channels(self.plans_manager, self.configuration_manager,
                                                               self.dataset_json)
        self.network = self.build_network_architecture(self.plans_manager, self.dataset_json,
                                                       self.configuration_manager, self.num_input_channels,
                                                       enable_deep_supervision=True).to(self.device)
        self._sync_replicas()
        self.optimizer, self.lr_scheduler = self.configure_optimizers()
        self._setup_grad_arenas()
        self.loss = self._build_loss()
        self.was_initialized = True

    def _networks(self) -> List[nn.Module]:
        return [self.network]

    def _sync_replicas(self):
        """every rank starts from rank 0's weights, as the DistributedDataParallel constructor guarantees in the
        reference (MVDTrainer.py:236-238; run_training never seeds, so each rank's He initialisation differs)."""
        if self.is_ddp:
            broadcast_parameters(self._networks(), src=0)

    def _setup_grad_arenas(self):
        """replaces DDP(self.network, device_ids=[local_rank]) (nnUNetTrainer.py:236-238): gradients live in flat
        arenas, buckets are all-reduced as backward produces them."""
        self._arenas = [GradArena(list(n.parameters()), world_size=None if self.is_ddp else 1) for n in self._networks()]
        lookup = {}
        for a in self._arenas:
            for p in a.params:
                lookup[id(p)] = a

        def alloc(p):
            a = lookup.get(id(p))
            return None if a is None else a.view_for(p)

        def ready(params):
            for p in params:
                a = lookup.get(id(p))
                if a is not None:
                    a.on_params_ready([p])

        ops.set_grad_allocator(alloc)
        ops.clear_param_grad_ready_hooks()
        ops.add_param_grad_ready_hook(ready)
        # gradients the kernels ACCUMULATE into (conv biases in front of InstanceNorm, head weights / biases) and those
        # of a head that may receive no gradient at all (zero-weighted scale): cleared by one launch per arena at the
        # top of every step instead of a fill per tensor
        views = []
        for net, a in zip(self._networks(), self._arenas):
            acc_params = [m.bias for m in net.modules() if isinstance(m, nn.Conv3d) and m.bias is not None]
            acc_params += [m.weight for m in net.decoder.seg_layers]
            views += a.prezero(acc_params)
        ops.set_prezeroed(views)
        # conv weights: optimiser update and refresh of their packed bf16 layouts in one kernel (no pack pass per step)
        if hasattr(self.optimizer, 'attach_weight_packers'):
            self.optimizer.attach_weight_packers([n._weight_packer() for n in self._networks()
                                                  if hasattr(n, '_weight_packer')])

    @staticmethod
    def build_network_architecture(plans_manager, dataset_json, configuration_manager, num_input_channels,
                                   enable_deep_supervision: bool = True) -> nn.Module:
        return get_network_from_plans(plans_manager, dataset_json, configuration_manager, num_input_channels,
                                      deep_supervision=enable_deep_supervision)

    def _get_deep_supervision_scales(self):
        """nnUNetTrainer.py:296-302: one scale per supervised output, None when deep supervision is off."""
        if not self.enable_deep_supervision:
            return None
        return list(list(i) for i in 1 / np.cumprod(np.vstack(self.configuration_manager.pool_op_kernel_sizes),
                                                     axis=0))[:-1]

    def _build_loss(self):
        if self.label_manager.has_regions:
            raise NotImplementedError('region-based training (DC_and_BCE_loss) is outside the built hot path')
        loss = DC_and_CE_loss({'batch_dice': self.configuration_manager.batch_dice, 'smooth': 1e-5, 'do_bg': False,
                               'ddp': self.is_ddp}, {}, weight_ce=1, weight_dice=1,
                              ignore_label=self.label_manager.ignore_label, dice_class=MemoryEfficientSoftDiceLoss)
        if self.enable_deep_supervision:
            deep_supervision_scales = self._get_deep_supervision_scales()
            weights = np.array([1 / (2 ** i) for i in range(len(deep_supervision_scales))])
            weights[-1] = 0
            weights = weights / weights.sum()
            loss = DeepSupervisionWrapper(loss, weights)
        return loss

    def configure_optimizers(self):
        params = [p for n in self._networks() for p in n.parameters()]
        optimizer = SGDNesterovClip(params, self.initial_lr, weight_decay=self.weight_decay, momentum=0.99,
                                    nesterov=True, max_norm=12.0)
        lr_scheduler = PolyLRScheduler(optimizer, self.initial_lr, self.num_epochs)
        return optimizer, lr_scheduler

    def set_deep_supervision_enabled(self, enabled: bool):
        for n in self._networks():
            n.decoder.deep_supervision = enabled

    def on_train_epoch_start(self):
        for n in self._networks():
            n.train()
        self.lr_scheduler.step(self.current_epoch)
        self.logger.log('lrs', self.optimizer.param_groups[0]['lr'], self.current_epoch)

    # ---- epoch-level hooks of the reference loop (MVDTrainer.py:808-1127), minus file / plot / dataloader management
    def on_train_start(self):
        if not self.was_initialized:
            self.initialize()
        self.set_deep_supervision_enabled(True)
        if self.dataloader_train is None:
            raise RuntimeError('run_training: set trainer.dataloader_train (and dataloader_val) to iterators of batch '
                               'dicts first; data loading / augmentation is outside this package')
        if self.output_folder is not None:
            os.makedirs(self.output_folder, exist_ok=True)
        if self.is_ddp:
            dist.barrier()

    def on_train_end(self):
        if self.output_folder is not None:
            self.save_checkpoint(os.path.join(self.output_folder, 'checkpoint_final.pth'))
            latest = os.path.join(self.output_folder, 'checkpoint_latest.pth')
            if self.local_rank == 0 and os.path.isfile(latest):
                os.remove(latest)

    def on_epoch_start(self):
        self.logger.log('epoch_start_timestamps', time.time(), self.current_epoch)

    def on_train_epoch_end(self, train_outputs: List[dict]):
        self.logger.log('train_losses', _mean_over_ranks(collate_outputs(train_outputs)['loss'], self.is_ddp), self.current_epoch)

    def on_validation_epoch_start(self):
        for n in self._networks():
            n.eval()

    def on_validation_epoch_end(self, val_outputs: List[dict]):
        """pseudo-Dice of the epoch from the hard tp / fp / fn counts summed over steps and ranks, 2tp / (2tp + fp + fn)
        per foreground class (MVDTrainer.py:1065-1096); a class that never occurs gives nan and is left out of the mean."""
        out = collate_outputs(val_outputs)
        tp, fp, fn = (_sum_over_ranks(np.sum(out[k], axis=0), self.is_ddp) for k in ('tp_hard', 'fp_hard', 'fn_hard'))
        with np.errstate(divide='ignore', invalid='ignore'):
            dice = [float(v) for v in 2 * tp / (2 * tp + fp + fn)]
        e = self.current_epoch
        self.logger.log('mean_fg_dice', np.nanmean(dice), e)
        self.logger.log('dice_per_class_or_region', dice, e)
        self.logger.log('val_losses', _mean_over_ranks(out['loss'], self.is_ddp), e)

    def on_epoch_end(self):
        """checkpoint cadence of MVDTrainer.py:1101-1127: 'latest' every save_every epochs, 'best' on a new EMA maximum."""
        e = self.current_epoch
        self.logger.log('epoch_end_timestamps', time.time(), e)
        folder = self.output_folder
        if folder is not None and (e + 1) % self.save_every == 0 and e != self.num_epochs - 1:
            self.save_checkpoint(os.path.join(folder, 'checkpoint_latest.pth'))
        ema = self.logger.my_fantastic_logging['ema_fg_dice']
        if ema and (self._best_ema is None or ema[-1] > self._best_ema):
            self._best_ema = ema[-1]
            if folder is not None:
                self.save_checkpoint(os.path.join(folder, 'checkpoint_best.pth'))
        self.current_epoch = e + 1

    def run_training(self):
        """MVDTrainer.py:1323-1345, with the next training batch uploaded underneath the current step."""
        self.on_train_start()
        for epoch in range(self.current_epoch, self.num_epochs):
            self.on_epoch_start()
            self.on_train_epoch_start()
            train_outputs = []
            batches = (next(self.dataloader_train) for _ in range(self.num_iterations_per_epoch))
            for batch in self.prefetching(batches):
                train_outputs.append(self.train_step(batch))
            self.on_train_epoch_end(train_outputs)
            if self.dataloader_val is not None and self.num_val_iterations_per_epoch > 0:
                with torch.no_grad():
                    self.on_validation_epoch_start()
                    val_outputs = []
                    for batch_id in range(self.num_val_iterations_per_epoch):
                        val_outputs.append(self.validation_step(next(self.dataloader_val)))
                    self.on_validation_epoch_end(val_outputs)
            self.on_epoch_end()
        self.on_train_end()

    # ------------------------------------------------------------------------------------------------------------
    def _forward(self, data):
        """network forward only (everything the step can do before it has the targets)."""
        return self.network(data)

    def _loss(self, output, target):
        return self.loss(output, target), output

    def _forward_loss(self, data, target):
        return self._loss(self._forward(data), target)

    def _to_device(self, batch):
        data = batch['data'].to(self.device, non_blocking=True)
        target = batch['target']
        if isinstance(target, list):
            target = [i.to(self.device, non_blocking=True) for i in target]
        else:
            target = target.to(self.device, non_blocking=True)
            if self.enable_deep_supervision:
                # a batch that carries only the full-resolution segmentation: the coarser deep-supervision targets are
                # produced on the GPU (the reference's CPU workers run DownsampleSegForDSTransform2, MVDTrainer.py:757-760)
                target = downsample_seg_for_ds(target, self._get_deep_supervision_scales())
        return data, target

    def train_step(self, batch: dict) -> dict:
        l = self.train_step_async(batch)
        st = getattr(self, '_graph_state', None)
        if self.use_cuda_graph and st is not None and st.get('graph2') is not None and l is st['loss'] and \
                os.environ.get('MVD_LOOKAHEAD', '1') != '0':
            # The reference reads the loss back every step (`l.detach().cpu().numpy()`, nnUNetTrainer.py:925), which leaves
            # the GPU idle from the end of this step until the host has returned and launched the next one.  With
            # `prefetching` the next batch is already staged: its FORWARD graph (weights of this step's update, no
            # dependence on anything the host still has to do) is queued first, and the loss comes back over the copy
            # stream behind this step's completion event -- the host wakes up when THIS step is done while the GPU is
            # already in the next forward pass.
            self._prelaunch_forward(st)
            side = st['copy_stream']
            if 'loss_host' not in st:
                st['loss_host'] = torch.empty(st['loss'].shape, dtype=st['loss'].dtype).pin_memory()
                st['ev_loss'] = torch.cuda.Event()
            side.wait_event(st['ev_done'])
            with torch.cuda.stream(side):
                st['loss_host'].copy_(st['loss'], non_blocking=True)
                st['ev_loss'].record(side)
            st['ev_loss'].synchronize()
            return {'loss': st['loss_host'].numpy().copy()}
        return {'loss': l.detach().cpu().numpy()}

    def _prelaunch_forward(self, st) -> None:
        """queue the forward graph of the batch `prefetching` has staged next (if it fits the captured shapes)"""
        up = getattr(self, '_upload_state', None)
        if up is None or len(up['staged']) != 1 or st.get('pre') is not None:
            return
        slot = next(iter(up['staged'].values()))
        if tuple(slot['data'].shape) != tuple(st['data'].shape) or slot['data'].dtype != st['data'].dtype:
            return
        main = torch.cuda.current_stream()
        main.wait_event(slot['ready'])
        st['data'].copy_(slot['data'], non_blocking=True)
        st['graph'].replay()
        st['pre'] = slot
        self._lookahead_launches = getattr(self, '_lookahead_launches', 0) + 1

    def _step_forward(self, data):
        self.optimizer.zero_grad(set_to_none=True)
        ops.begin_step(data.device)
        for a in self._arenas:
            a.begin_step()
        return self._forward(data)

    def _unit_gradient(self, loss: torch.Tensor) -> torch.Tensor:
        """a persistent tensor of ones shaped like the loss (created once per device / dtype: CUDA-graph friendly)"""
        key = (loss.device, loss.dtype, tuple(loss.shape))
        cache = self.__dict__.setdefault('_unit_gradients', {})
        one = cache.get(key)
        if one is None:
            one = cache[key] = torch.ones(loss.shape, dtype=loss.dtype, device=loss.device)
        return one

    def _step_body(self, data, target) -> torch.Tensor:
        return self._step_backward(self._step_forward(data), target)

    def _step_backward(self, output, target) -> torch.Tensor:
        l, _ = self._loss(output, target)
        l.backward(self._unit_gradient(l))      # explicit seed: autograd would launch a fill kernel for its implicit ones
        world = 1
        for a in self._arenas:
            a.finish()
            a.attach_grads()
            world = a.world_size
        self.optimizer.step(grad_scale=1.0 / world)
        return l.detach()

    # ---- one-batch-ahead upload ---------------------------------------------------------------------------------
    def prefetching(self, batches):
        """iterate host batches with the NEXT batch's host->device copy already in flight.

        ``for batch in trainer.prefetching(loader): trainer.train_step(batch)`` is the training loop of
        ``run_training`` (MVDTrainer.py:1323-1331: ``self.train_step(next(self.dataloader_train))``) with the upload of
        batch i+1 issued on a copy stream before batch i is stepped, so the PCIe transfer (50 MB per cfg-2 batch,
        ~0.6 ms) runs underneath the previous step's kernels instead of in front of its own forward pass.  The yielded
        objects are the host batches themselves; ``train_step`` recognises them and takes the staged device copies."""
        it = iter(batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        self._stage_upload(nxt)
        while nxt is not None:
            cur = nxt
            try:
                nxt = next(it)
            except StopIteration:
                nxt = None
            if nxt is not None:
                self._stage_upload(nxt)
            yield cur

    def _stage_upload(self, batch: dict) -> None:
        st = getattr(self, '_upload_state', None)
        if st is None:      # one copy stream and two staging slots per trainer, created on first use
            st = self._upload_state = dict(stream=torch.cuda.Stream(device=self.device), slots=[None, None], turn=0,
                                           staged={})
        k = st['turn']
        st['turn'] ^= 1
        target = batch['target'] if isinstance(batch['target'], list) else [batch['target']]
        slot = st['slots'][k]
        shapes = (tuple(batch['data'].shape), tuple(tuple(t.shape) for t in target))
        if slot is None or slot['shapes'] != shapes:
            slot = st['slots'][k] = dict(shapes=shapes,
                                         data=torch.empty(batch['data'].shape, dtype=batch['data'].dtype, device=self.device),
                                         target=[torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in target],
                                         ready=torch.cuda.Event(), consumed=None)
        side = st['stream']
        if slot['consumed'] is not None:
            side.wait_event(slot['consumed'])     # the step that used this slot has copied it into the graph inputs
        with torch.cuda.stream(side):
            slot['data'].copy_(batch['data'], non_blocking=True)
            for a, b in zip(slot['target'], target):
                a.copy_(b, non_blocking=True)
            slot['ready'].record(side)
        for key in [i for i, v in st['staged'].items() if v is slot]:
            del st['staged'][key]
        st['staged'][id(batch)] = slot

    def _take_staged(self, batch: dict):
        st = getattr(self, '_upload_state', None)
        slot = st['staged'].pop(id(batch), None) if st else None
        if slot is None:
            return batch, None
        torch.cuda.current_stream().wait_event(slot['ready'])
        tgt = slot['target'] if isinstance(batch['target'], list) else slot['target'][0]
        return {'data': slot['data'], 'target': tgt}, slot

    def train_step_async(self, batch: dict) -> torch.Tensor:
        batch, slot = self._take_staged(batch)
        loss = self._train_step_async(batch)
        if slot is not None:
            # inputs have been copied into the step's own buffers (or consumed by its kernels) in stream order
            slot['consumed'] = torch.cuda.Event()
            slot['consumed'].record()
        return loss

    def _train_step_async(self, batch: dict) -> torch.Tensor:
        """the step without the device->host read of the loss (returned as a device scalar).

        With ``self.use_cuda_graph`` the whole step (forward, losses, backward, gradient exchange, clip + SGD: a few
        hundred launches) is captured once into a CUDA graph and replayed, which removes the launch gaps between the
        many short kernels.  The first ``graph_warmup_steps`` steps run eagerly (lazy one-time initialisation must not
        happen under capture); the graph is re-captured when the batch shapes or the learning rate (a kernel
        argument, changed once per epoch by PolyLR) change."""
        st = getattr(self, '_graph_state', None)
        if self.use_cuda_graph and st is not None:
            # steady state: copy the host batch straight into the graph's static input buffers (no staging tensors)
            data, target = batch['data'], batch['target']
            if not isinstance(target, list):
                target = [target]
            key = (tuple(data.shape), tuple(tuple(t.shape) for t in target),
                   tuple(g['lr'] for g in self.optimizer.param_groups))
            if st['key'] == key and data.dtype == st['data'].dtype and all(a.dtype == b.dtype for a, b in zip(st['target'], target)):
                return self._replay(st, data, target)
        data, target = self._to_device(batch)
        if not self.use_cuda_graph:
            return self._step_body(data, target)
        if not isinstance(target, list):
            target = [target]
        self._eager_steps_done = getattr(self, '_eager_steps_done', 0)
        if self._eager_steps_done < self.graph_warmup_steps:
            self._eager_steps_done += 1
            return self._step_body(data, target)
        key = (tuple(data.shape), tuple(tuple(t.shape) for t in target),
               tuple(g['lr'] for g in self.optimizer.param_groups))
        if st is None or st['key'] != key:
            self._graph_state = None
            self.optimizer.ensure_state()    # lazy momentum initialisation must not end up inside the graph
            sd = data.clone()
            stg = [t.clone() for t in target]
            torch.cuda.synchronize()
            if self.split_graph:
                # two graphs sharing one memory pool: the forward pass only needs `data`, so the host->device copy of
                # the targets (36 % of the batch bytes) can run on a copy stream underneath it
                pool = torch.cuda.graph_pool_handle()
                g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1, pool=pool, stream=self._capture_stream()):
                    outs = self._step_forward(sd)
                with torch.cuda.graph(g2, pool=pool, stream=self._capture_stream()):
                    loss = self._step_backward(outs, stg)
                st = self._graph_state = dict(key=key, graph=g1, graph2=g2, data=sd, target=stg, loss=loss, outs=outs,
                                              copy_stream=torch.cuda.Stream(device=sd.device),
                                              ev_targets=torch.cuda.Event(), ev_done=torch.cuda.Event())
                st['ev_done'].record()
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._capture_stream()):
                    loss = self._step_body(sd, stg)
                st = self._graph_state = dict(key=key, graph=g, graph2=None, data=sd, target=stg, loss=loss)
            st['graph'].replay()
            if st['graph2'] is not None:
                st['graph2'].replay()
                st['ev_done'].record()
            return st['loss']
        return self._replay(st, data, target)

    def _capture_stream(self):
        """the step is captured on a HIGH-priority stream: its kernel nodes (the dependency chain of the step) are
        scheduled ahead of the deferred weight-gradient kernels, which run on a lowest-priority side stream and fill the
        SMs the chain leaves free.  MVD_STREAM_PRIORITY=0 captures on a normal stream."""
        if os.environ.get('MVD_STREAM_PRIORITY', '1') == '0':
            return None
        s = getattr(self, '_hp_stream', None)
        if s is None:
            s = self._hp_stream = torch.cuda.Stream(device=self.device, priority=-1)
        return s

    def _replay(self, st, data, target) -> torch.Tensor:
        """copy a batch (host or device tensors) into the graph's static inputs and replay."""
        if st['graph2'] is None:
            st['data'].copy_(data, non_blocking=True)
            for a, b in zip(st['target'], target):
                a.copy_(b, non_blocking=True)
            st['graph'].replay()
            return st['loss']
        main = torch.cuda.current_stream()
        side = st['copy_stream']
        pre = st.pop('pre', None)
        launched = pre is not None and data is pre['data']      # train_step queued this batch's forward pass already
        if not launched:
            st['data'].copy_(data, non_blocking=True)
        side.wait_event(st['ev_done'])              # the previous step has finished reading the static targets
        if target and target[0].is_cuda:             # staged device batch: its upload is ordered before `main` only
            ev = st.setdefault('ev_start', torch.cuda.Event())
            ev.record(main)
            side.wait_event(ev)
        with torch.cuda.stream(side):
            for a, b in zip(st['target'], target):
                a.copy_(b, non_blocking=True)
            st['ev_targets'].record(side)
        if not launched:
            st['graph'].replay()                     # forward: needs `data` only
        main.wait_event(st['ev_targets'])
        st['graph2'].replay()                        # losses, backward, exchange, optimiser
        st['ev_done'].record(main)
        return st['loss']

    def validation_step(self, batch: dict) -> dict:
        data, target = self._to_device(batch)
        with torch.no_grad():
            l, output = self._forward_loss(data, target)
        if self.enable_deep_supervision:
            output, target = output[0], target[0]
        tp, fp, fn = ops.argmax_tp_fp_fn(output, target)
        tp_hard, fp_hard, fn_hard = (t.detach().cpu().numpy()[1:] for t in (tp, fp, fn))
        return {'loss': l.detach().cpu().numpy(), 'tp_hard': tp_hard, 'fp_hard': fp_hard, 'fn_hard': fn_hard}

    # ------------------------------------------------------------------------------------------------------------
    def _checkpoint_dict(self) -> dict:
        """keys of MVDTrainer.py:1129-1152 (subclasses extend the dict)."""
        return {
            'network_weights': self.network.state_dict(),
            'optimizer_state': self.optimizer.state_dict(),
            'grad_scaler_state': None,
            'logging': self.logger.get_checkpoint(),
            '_best_ema': self._best_ema,
            'current_epoch': self.current_epoch + 1,
            'init_args': {'configuration': self.configuration_name, 'fold': self.fold},
            'trainer_name': self.__class__.__name__,
            'inference_allowed_mirroring_axes': None,
        }

    def save_checkpoint(self, filename: str) -> None:
        if self.local_rank != 0:
            return
        tmp = filename + '.tmp'
        torch.save(self._checkpoint_dict(), tmp)      # written once, published atomically
        os.replace(tmp, filename)

    @staticmethod
    def _strip_ddp_prefix(weights: dict, own_keys) -> dict:
        """a checkpoint saved from a DDP-wrapped module carries 'module.' in front of every key (MVDTrainer.py:1162-1167)."""
        return {(k[7:] if k not in own_keys and k.startswith('module.') else k): v for k, v in weights.items()}

    def _load_networks(self, ck: dict) -> None:
        self.network.load_state_dict(self._strip_ddp_prefix(ck['network_weights'], self.network.state_dict().keys()))

    def load_checkpoint(self, filename_or_checkpoint: Union[dict, str]) -> None:
        """MVDTrainer.py:1154-1190."""
        if not self.was_initialized:
            self.initialize()
        ck = filename_or_checkpoint
        if isinstance(ck, str):
            ck = torch.load(ck, map_location=self.device, weights_only=False)
        self._load_networks(ck)
        self.optimizer.load_state_dict(ck['optimizer_state'])
        self.current_epoch = ck['current_epoch']
        if ck.get('logging'):
            self.logger.load_checkpoint(ck['logging'])
        self._best_ema = ck.get('_best_ema')
        self._sync_replicas()
        # a captured step holds the addresses of the momentum buffers that load_state_dict just replaced
        self._graph_state = None


class MVDTrainer(nnUNetTrainer):
    """canonical mutual-distillation step (the reference's ContrastiveTrainer.train_step, MVDTrainer.py:879-985;
    decisions of SURVEY.md section 8c): net1 sees modality 0, net2 modality 1 (split selfattnNet.py:588-589);
    total = L(out1,tgt) + L(out2,tgt) + lambda3 * topo + lambda1 * KL  (MVDTrainer.py:925, lambdas :132-134).
    The memory-bank contrastive branch (:927-972) depends on modules missing from the reference tree and is not built."""

    def __init__(self, *a, topo_iter: Optional[int] = 3, kl_T: float = 1.0, kl_vessel_only: bool = False, **kw):
        super().__init__(*a, **kw)
        self.lambda1, self.lambda2, self.lambda3 = 0.5, 0.1, 1.0
        self.vessel_class = 2
        self.topo_iter = topo_iter
        self.kl_T = kl_T
        self.kl_vessel_only = kl_vessel_only
        self.network2 = None
        # the two modality networks are independent until the loss: run network 2 on a second stream so that its
        # latency-bound low-resolution layers (grids smaller than the machine) and its HBM-bound passes fill in under
        # network 1's kernels.  MVD_CONCURRENT_NETS=0 runs them back to back on one stream.
        self.concurrent_networks = os.environ.get('MVD_CONCURRENT_NETS', '1') != '0'
        self._net2_stream = None

    def initialize(self):
        if self.was_initialized:
            raise RuntimeError('You have called self.initialize even though the trainer was already initialized.')
        n_mod = determine_num_input_channels(self.plans_manager, self.configuration_manager, self.dataset_json)
        assert n_mod == 2, 'the mutual-distillation step needs exactly two modalities'
        self.num_input_channels = n_mod
        self.network = self.build_network_architecture(self.plans_manager, self.dataset_json,
                                                       self.configuration_manager, 1, True).to(self.device)
        self.network2 = self.build_network_architecture(self.plans_manager, self.dataset_json,
                                                        self.configuration_manager, 1, True).to(self.device)
        self._sync_replicas()
        self.optimizer, self.lr_scheduler = self.configure_optimizers()
        self._setup_grad_arenas()
        self.loss = self._build_loss()
        self.topo = soft_cldice(iter_=self.topo_iter, smooth=1.) if self.topo_iter is not None else None
        self.was_initialized = True

    def _networks(self):
        return [self.network, self.network2]

    def _forward(self, data):
        if not self.concurrent_networks:
            return self.network(data[:, 0:1]), self.network2(data[:, 1:2])
        main = torch.cuda.current_stream()
        if self._net2_stream is None:
            self._net2_stream = torch.cuda.Stream(device=self.device, priority=-1)
        side = self._net2_stream
        side.wait_stream(main)                 # input, zero pool and weights are ready on the compute stream
        with torch.cuda.stream(side):
            out2 = self.network2(data[:, 1:2])
        out1 = self.network(data[:, 0:1])
        main.wait_stream(side)
        return out1, out2

    def _step_backward(self, output, target) -> torch.Tensor:
        if not self.concurrent_networks or self._net2_stream is None:
            return super()._step_backward(output, target)
        l, _ = self._loss(output, target)
        l.backward(self._unit_gradient(l))      # explicit seed: autograd would launch a fill kernel for its implicit ones
        # network 2's backward ran on its own stream (autograd keeps every node on its forward stream); parameter
        # gradients bypass AccumulateGrad here, so the engine has no leaf stream to join: do it explicitly
        torch.cuda.current_stream().wait_stream(self._net2_stream)
        world = 1
        for a in self._arenas:
            a.finish()
            a.attach_grads()
            world = a.world_size
        self.optimizer.step(grad_scale=1.0 / world)
        return l.detach()

    def _loss(self, output, target):
        out1, out2 = output
        if hasattr(self.loss, 'forward_networks'):      # loss(out1, tgt) + loss(out2, tgt) through one fused node
            l = self.loss.forward_networks([out1, out2], target)
        else:
            l = self.loss(out1, target) + self.loss(out2, target)
        c = self.vessel_class
        if self.kl_vessel_only:
            mutual = distill_kl(out1[0][:, c:c + 1], out2[0][:, c:c + 1], self.kl_T)
        else:
            mutual = distill_kl(out1[0], out2[0], self.kl_T, upstream_grad=self.lambda1)
        l = l + self.lambda1 * mutual
        if self.topo is not None:
            prob, gt = softmax_channel(out1[0], c, target[0])      # softmax channel c and (target == c), one launch
            l = l + self.lambda3 * self.topo(gt, prob)
        self.last_terms = dict(mutual=mutual.detach())
        return l, out1

    def _checkpoint_dict(self) -> dict:
        ck = super()._checkpoint_dict()
        ck['network2_weights'] = self.network2.state_dict()
        return ck

    def _load_networks(self, ck: dict) -> None:
        if 'network2_weights' not in ck:
            raise KeyError("checkpoint has no 'network2_weights': it was not written by MVDTrainer (a single-network or "
                           'reference-format checkpoint would leave the second modality network at its random '
                           'initialisation while restoring optimiser momentum for it); load it into nnUNetTrainer, or '
                           'add the second network\'s state_dict under that key')
        super()._load_networks(ck)
        self.network2.load_state_dict(self._strip_ddp_prefix(ck['network2_weights'], self.network2.state_dict().keys()))
