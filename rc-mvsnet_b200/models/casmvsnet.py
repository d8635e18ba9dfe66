"""Drop-in for the reference's models/casmvsnet.py (CascadeMVSNet / CascadeMVSNet_eval / DepthNet / DepthNet_eval).

Same constructor arguments, forward() signatures, output dictionaries and state_dict keys as the reference
(models/casmvsnet.py:45-124, 126-231, 234-311, 313-417); the per-stage work runs on librcmvs.so kernels:

    hypotheses (a5) -> fused plane sweep + variance (a1+a2) -> CostRegNet (a3) -> softmax/soft-argmin/conf (a4)

with channels-last volumes in between, no [B,D,H,W] full-resolution hypothesis tensor, no warped-volume
temporaries, no torch.inverse and no host synchronisation anywhere in forward().
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from rcmvs import ops
from rcmvs.strict import library_fallback
from .modules import CostRegNet, FeatureNet, as_channels_last_views, homo_warping

Align_Corners_Range = False
_STAGE_SCALE = {"stage1": 4, "stage2": 2, "stage3": 1}


def _stack_views_cl(features):
    """list of V tensors [B,C,h,w] -> [B,V,h,w,C] contiguous (one kernel per view only if not already NHWC)."""
    return torch.stack([as_channels_last_views(f) for f in features], dim=1)


def _sweep_regularise_regress(feats_cl, proj_matrices, depth_values, cost_regularization, want_noref=False):
    """feats_cl [B,V,h,w,C]; proj_matrices [B,V,2,4,4]; depth_values [B,D] (uniform planes) or [B,D,h,w]."""
    rt = ops.relative_projections(proj_matrices.float().contiguous())
    depth_values = depth_values.float().contiguous()
    if not want_noref and cost_regularization.wants_split_input() and min(feats_cl.shape[2:4]) >= 2:
        var, noref = ops.costvolume_variance_split(feats_cl, rt, depth_values), None      # bf16 hi/lo planes for the TMA conv
    else:
        res = ops.costvolume_variance(feats_cl, rt, depth_values, want_noref)
        var, noref = res if want_noref else (res, None)
    cost = cost_regularization.forward_cl(var)
    depth, conf = ops.softargmin_conf(cost, depth_values)
    return depth, conf, noref, rt


class DepthNet_eval(nn.Module):
    """reference models/casmvsnet.py:234-311."""

    def forward(self, features, proj_matrices, depth_values, num_depth, cost_regularization, imgs, pad=0,
                prob_volume_init=None):
        assert len(features) == proj_matrices.shape[1], "Different number of images and projection matrices"
        assert depth_values.shape[1] == num_depth, "depth_values.shape[1]:{}  num_depth:{}".format(
            depth_values.shape[1], num_depth)
        if prob_volume_init is not None:
            raise NotImplementedError("prob_volume_init is never passed by the reference callers")
        if cost_regularization.needs_autograd() or (torch.is_grad_enabled() and any(f.requires_grad for f in features)):
            out = DepthNet._forward_autograd(self, features, proj_matrices, depth_values, num_depth, cost_regularization,
                                             imgs)
            return {"depth": out["depth"], "photometric_confidence": out["photometric_confidence"]}
        with torch.no_grad():
            depth, conf, _, _ = _sweep_regularise_regress(_stack_views_cl(features), proj_matrices, depth_values,
                                                          cost_regularization)
        return {"depth": depth, "photometric_confidence": conf}


class DepthNet(nn.Module):
    """reference models/casmvsnet.py:45-124 (training class: also returns volume_feature_no_ref =
    cat[warped source RGB at stage resolution (3·(V-1) ch), variance over the sources only (C ch)])."""

    def forward(self, features, proj_matrices, depth_values, num_depth, cost_regularization, imgs, pad=0,
                prob_volume_init=None):
        assert len(features) == proj_matrices.shape[1], "Different number of images and projection matrices"
        assert depth_values.shape[1] == num_depth, "depth_values.shape[1]:{}  num_depth:{}".format(
            depth_values.shape[1], num_depth)
        if prob_volume_init is not None:
            raise NotImplementedError("prob_volume_init is never passed by the reference callers")
        needs_grad = torch.is_grad_enabled() and (any(f.requires_grad for f in features) or
                                                  cost_regularization.needs_autograd())
        if needs_grad or cost_regularization.training:
            return self._forward_autograd(features, proj_matrices, depth_values, num_depth, cost_regularization, imgs)
        B, V = imgs.shape[:2]
        C, h, w = features[0].shape[1:]
        with torch.no_grad():
            depth, conf, noref, rt = _sweep_regularise_regress(_stack_views_cl(features), proj_matrices,
                                                              depth_values, cost_regularization, want_noref=True)
            small = F.interpolate(imgs.reshape(B * V, *imgs.shape[2:]).float(), (h, w), mode="bilinear",
                                  align_corners=False).view(B, V, 3, h, w)
            dv = depth_values.float().contiguous()
            vol = torch.empty((B, 3 * (V - 1) + C, num_depth, h, w), device=imgs.device, dtype=torch.float32)
            for i in range(V - 1):
                vol[:, 3 * i:3 * i + 3] = ops.homo_warp_fwd(small[:, i + 1].contiguous(), rt[:, i].contiguous(), dv)
            vol[:, 3 * (V - 1):] = ops.to_planar(noref)
        return {"depth": depth, "photometric_confidence": conf, "volume_feature_no_ref": vol}

    def _forward_autograd(self, features, proj_matrices, depth_values, num_depth, cost_regularization, imgs):
        """Differentiable path for train_rcmvsnet.py, all on own kernels (rcmvs.autograd): fused plane sweep + variance
        (+ var_no_ref) with its hand-written backward into the V feature maps, CostRegNet with batch-statistics
        BatchNorm, softmax + depth regression with its backward.  The hypotheses carry no gradient (the sample grid is
        built under no_grad in the reference, modules.py:314, and grad_method='detach')."""
        from rcmvs import autograd as A
        B, V = imgs.shape[:2]
        C, h, w = features[0].shape[1:]
        with torch.no_grad():
            rt = ops.relative_projections(proj_matrices.float().contiguous())
            dv = depth_values.detach().float().contiguous()
            small = F.interpolate(imgs.reshape(B * V, *imgs.shape[2:]).float(), (h, w), mode="bilinear",
                                  align_corners=False).view(B, V, 3, h, w)
            rgb = [ops.homo_warp_fwd(small[:, i + 1].contiguous(), rt[:, i].contiguous(), dv) for i in range(V - 1)]
        feats_cl = torch.stack([as_channels_last_views(f) for f in features], dim=1)        # [B,V,h,w,C]
        var, noref = A.costvolume_variance(feats_cl, rt, dv, True)
        cost = cost_regularization.forward_cl_train(var)
        depth, conf = A.softargmin(cost, dv)
        return {"depth": depth, "photometric_confidence": conf,
                "volume_feature_no_ref": torch.cat(rgb + [A.to_planar(noref)], dim=1)}


class _CascadeBase(nn.Module):
    depthnet_cls = None

    def __init__(self, refine=False, ndepths=[48, 32, 8], depth_interals_ratio=[4, 2, 1], share_cr=False,
                 grad_method="detach", arch_mode="fpn", cr_base_chs=[8, 8, 8]):
        super().__init__()
        if refine:
            raise NotImplementedError("refine=True is dead code in the reference (RefineNet uses F.cat)")
        assert len(ndepths) == len(depth_interals_ratio)
        self.refine, self.share_cr, self.ndepths = refine, share_cr, list(ndepths)
        self.depth_interals_ratio, self.grad_method = list(depth_interals_ratio), grad_method
        self.arch_mode, self.cr_base_chs, self.num_stage = arch_mode, list(cr_base_chs), len(ndepths)
        self.stage_infos = {k: {"scale": float(v)} for k, v in _STAGE_SCALE.items()}
        self.feature = FeatureNet(base_channels=8, stride=4, num_stage=self.num_stage, arch_mode=arch_mode)
        if share_cr:
            self.cost_regularization = CostRegNet(in_channels=self.feature.out_channels, base_channels=8)
        else:
            self.cost_regularization = nn.ModuleList(
                [CostRegNet(in_channels=self.feature.out_channels[i], base_channels=self.cr_base_chs[i])
                 for i in range(self.num_stage)])
        self.DepthNet = self.depthnet_cls()

    def _cr(self, i):
        return self.cost_regularization if self.share_cr else self.cost_regularization[i]

    def _features(self, imgs):
        """FeatureNet on all views.  Eval: one batched call, channels-last outputs [B,V,h,w,C] per stage.
        Train: one call per view like the reference (BatchNorm statistics are per call, casmvsnet.py:178-180)."""
        B, V = imgs.shape[:2]
        if self.training:
            per_view = [self.feature(imgs[:, v]) for v in range(V)]
            return per_view, None
        flat = imgs.reshape(B * V, *imgs.shape[2:])
        if flat.is_cuda and not torch.is_grad_enabled() and self.feature.base_channels == 8:
            # stage-2/3 feature maps are produced on a side stream while stage 1 runs (joined in _run)
            # (not while a CUDA graph is being captured: blocks freed inside a capture are reused in capture order, which the
            # cross-stream record_stream bookkeeping of the side-stream path cannot protect)
            side = None if torch.cuda.is_current_stream_capturing() else self._side_stream(flat.device)
            out = self.feature.forward_cl(flat, side_stream=side)
            stacked = {k: (v if k == "_ready" else v.view(B, V, *v.shape[1:])) for k, v in out.items()}
            return None, stacked
        out = self.feature(flat)
        stacked = {}
        for k, f in out.items():
            cl = as_channels_last_views(f)                     # [B*V,h,w,C]
            stacked[k] = cl.view(B, V, *cl.shape[1:])
        return None, stacked

    def _side_stream(self, device):
        streams = self.__dict__.setdefault("_side_streams", {})
        if device not in streams:
            streams[device] = torch.cuda.Stream(device)
        return streams[device]

    def _run(self, imgs, proj_matrices, depth_values):
        B, V, _, H, W = imgs.shape
        imgs = imgs.float()
        depth_values = depth_values.float().contiguous()
        if self.grad_method != "detach" and torch.is_grad_enabled():
            # 'undetach' (a train_rcmvsnet.py choice) differentiates the next stage's hypotheses w.r.t. this stage's depth;
            # the hypothesis kernels carry no gradient, so refuse instead of silently dropping that term
            raise NotImplementedError("rcmvs: grad_method='undetach' is not implemented on the CUDA path (use 'detach', "
                                      "the reference default)")
        per_view, stacked = self._features(imgs)
        feats_ready = stacked.pop("_ready", None) if stacked is not None else None
        fused = stacked is not None and (not torch.is_grad_enabled() or self.depthnet_cls is DepthNet_eval)
        if not fused and stacked is not None:
            per_view = [{k: stacked[k][:, v].permute(0, 3, 1, 2) for k in stacked} for v in range(V)]
        outputs, depth = {}, None
        for s in range(self.num_stage):
            key = "stage%d" % (s + 1)
            scale = _STAGE_SCALE[key]
            h, w, D = H // scale, W // scale, self.ndepths[s]
            if depth is None:
                planes = ops.depth_hypotheses_first(depth_values, D)               # [B,D]: uniform per plane
                hyp = planes
            else:
                if feats_ready is not None:                # join the FeatureNet tail before its maps are first read
                    torch.cuda.current_stream(imgs.device).wait_event(feats_ready)
                    feats_ready = None
                # depth_interval = (max - min) / Dv of batch element 0 is taken on the device (the reference pulls
                # min/max to the host, :172-174): no host synchronisation anywhere in the forward
                cur = depth.detach() if self.grad_method == "detach" else depth
                hyp = ops.depth_hypotheses_next_dev(cur.contiguous(), depth_values, D, self.depth_interals_ratio[s],
                                                    H, W, h, w)
            if fused:
                with torch.no_grad():
                    want_noref = self.depthnet_cls is DepthNet and s == 0
                    d, c, noref, rt = _sweep_regularise_regress(stacked[key], proj_matrices[key], hyp, self._cr(s),
                                                                want_noref)
                    st = {"depth": d, "photometric_confidence": c}
                    if want_noref:    # only stage 1's is consumed by the caller (reference casmvsnet.py:231)
                        st["volume_feature_no_ref"] = self._volume_feature(imgs, rt, hyp, noref, h, w)
            else:
                hyp4 = hyp if hyp.dim() == 4 else hyp.view(B, D, 1, 1).expand(-1, -1, h, w).contiguous()
                st = self.DepthNet([f[key] for f in per_view], proj_matrices[key], depth_values=hyp4, num_depth=D,
                                   cost_regularization=self._cr(s), imgs=imgs)
            depth = st["depth"]
            outputs[key] = st
            outputs.update(st)
        return outputs

    @staticmethod
    def _volume_feature(imgs, rt, hyp, noref, h, w):
        """cat[warped source RGB (3·(V-1)), var_no_ref (C)] as [B,3(V-1)+C,D,h,w] (casmvsnet.py:59-60,83,100-102)."""
        B, V = imgs.shape[:2]
        D = hyp.shape[1]
        small = F.interpolate(imgs.reshape(B * V, *imgs.shape[2:]), (h, w), mode="bilinear",
                              align_corners=False).view(B, V, 3, h, w)
        C = noref.shape[-1]
        vol = torch.empty((B, 3 * (V - 1) + C, D, h, w), device=imgs.device, dtype=torch.float32)
        for i in range(V - 1):
            vol[:, 3 * i:3 * i + 3] = ops.homo_warp_fwd(small[:, i + 1].contiguous(), rt[:, i].contiguous(), hyp)
        vol[:, 3 * (V - 1):] = ops.to_planar(noref)
        return vol


class CascadeMVSNet_eval(_CascadeBase):
    """reference models/casmvsnet.py:313-417: forward(imgs [B,V,3,H,W], proj_matrices {stageK: [B,V,2,4,4]},
    depth_values [B,Dv]) -> outputs dict."""
    depthnet_cls = DepthNet_eval

    def forward(self, imgs, proj_matrices, depth_values):
        return self._run(imgs, proj_matrices, depth_values)


class CascadeMVSNet(_CascadeBase):
    """reference models/casmvsnet.py:126-231: returns (outputs, outputs['stage1']['volume_feature_no_ref'])."""
    depthnet_cls = DepthNet

    def forward(self, imgs, proj_matrices, depth_values):
        outputs = self._run(imgs, proj_matrices, depth_values)
        return outputs, outputs["stage1"]["volume_feature_no_ref"]
