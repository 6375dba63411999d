"""DistributedStrategy. Parity: python/paddle/distributed/fleet/base/distributed_strategy.py (dygraph-relevant subset +
the static-mode switches kept as plain attributes so existing configs load)."""
from __future__ import annotations

import copy


class _Cfg(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class DistributedStrategy:
    def __init__(self):
        self.hybrid_configs = _Cfg(dp_degree=1, mp_degree=1, pp_degree=1, sharding_degree=1, sep_degree=1,
                                   order=["dp", "pp", "sharding", "sep", "mp"],
                                   mp_configs=_Cfg(sync_param=False, sync_grad=False, sync_moment=False, mp_async_allreduce=False,
                                                   mp_skip_c_identity=False, mp_fused_linear_param_grad_add=False),
                                   pp_configs=_Cfg(delay_scale_loss=False, dp_comm_overlap=False, sharding_comm_overlap=False,
                                                   enable_timer=False, release_gradients=False, overlap_p2p_comm=False,
                                                   clear_every_step_cache=False, use_batch_p2p_comm=True),
                                   sharding_configs=_Cfg(tensor_fusion=False, accumulate_steps=1, comm_overlap=False, split_param=False))
        self.pipeline_configs = _Cfg(accumulate_steps=1, micro_batch_size=1, schedule_mode="1F1B", p2p_cache_shape=True, enable_partial_send_recv=True)
        self.tensor_parallel_configs = _Cfg(tensor_parallel_degree=1, tensor_init_seed=-1)
        self.sharding_configs = _Cfg(sharding_degree=1, stage=1, segment_broadcast_MB=32.0)
        self.recompute_configs = _Cfg(checkpoints=[], enable_offload=False)
        self.amp_configs = _Cfg(init_loss_scaling=32768.0, use_pure_fp16=False, use_bf16=False, custom_white_list=[], custom_black_list=[])
        self.gradient_merge_configs = _Cfg(k_steps=1, avg=True)
        self.lamb_configs = _Cfg(lamb_weight_decay=0.01, exclude_from_weight_decay=[])
        self.a_sync_configs = _Cfg(k_steps=-1)
        self.amp = False
        self.recompute = False
        self.pipeline = False
        self.tensor_parallel = False
        self.sharding = False
        self.gradient_merge = False
        self.lamb = False
        self.lars = False
        self.dgc = False
        self.localsgd = False
        self.a_sync = False
        self.heter_ccl_mode = False
        self.find_unused_parameters = False
        self.fuse_all_reduce_ops = True
        self.fuse_grad_size_in_MB = 32
        self.fuse_grad_merge = False
        self.last_comm_group_size_MB = 1
        self.nccl_comm_num = 1
        self.sync_nccl_allreduce = True
        self.without_graph_optimization = True
        self.auto = False
        self.semi_auto = False
        self.auto_search = False
        self.split_data = True
        self.elastic = False
        self.asp = False
        self.qat = False
        self.qat_configs = _Cfg(channel_wise_abs_max=True, weight_bits=8, activation_bits=8, not_quant_pattern=[], algo=None)
        self.adaptive_localsgd = False
        self.localsgd_configs = _Cfg(k_steps=1, begin_step=1)
        self.adaptive_localsgd_configs = _Cfg(init_k_steps=1, begin_step=1)
        self.dgc_configs = _Cfg(rampup_begin_step=0, rampup_step=1, sparsity=[0.999])
        self.lars_configs = _Cfg(lars_coeff=0.001, lars_weight_decay=0.0005, epsilon=0.0, exclude_from_weight_decay=[])
        self.fp16_allreduce = False
        self.sync_batch_norm = False
        self.use_hierarchical_allreduce = False
        self.hierarchical_allreduce_inter_nranks = 1
        self.fuse_grad_size_in_num = 8
        self.adam_d2sum = False
        self.is_fl_ps_mode = False
        self.is_with_coordinator = False
        self.cudnn_exhaustive_search = False
        self.conv_workspace_size_limit = 512
        self.cudnn_batchnorm_spatial_persistent = False
        self.gradient_scale_configs = _Cfg(scale_strategy="avg", scale_gradient=False)
        self.trainer_desc_configs = _Cfg(dump_fields_path="", dump_fields=[], dump_param=[], stat_var_names=[])
        self.sparse_table_configs = _Cfg()
        self.fleet_desc_configs = _Cfg()
        self.fs_client_param = _Cfg(uri="", user="", passwd="", hadoop_bin="")
        self.build_strategy = None

    def __setattr__(self, k, v):
        if k.endswith("_configs") and isinstance(v, dict) and k in self.__dict__:
            cur = self.__dict__[k]
            for kk, vv in v.items():
                if isinstance(vv, dict) and isinstance(cur.get(kk), dict):
                    cur[kk].update(vv)
                else:
                    cur[kk] = vv
        else:
            object.__setattr__(self, k, _Cfg(v) if isinstance(v, dict) and not isinstance(v, _Cfg) else v)

    def __deepcopy__(self, memo):
        s = DistributedStrategy()
        for k, v in self.__dict__.items():
            object.__setattr__(s, k, copy.deepcopy(v, memo))
        return s

    def save_to_prototxt(self, path):
        import json

        with open(path, "w") as f:
            json.dump({k: v for k, v in self.__dict__.items()}, f, default=str, indent=1)

    def load_from_prototxt(self, path):
        import json

        with open(path) as f:
            for k, v in json.load(f).items():
                setattr(self, k, v)

    def __repr__(self):
        return "DistributedStrategy(" + ", ".join(f"{k}={v}" for k, v in self.hybrid_configs.items() if k.endswith("degree")) + ")"
