"""Drop-ins for the reference's `loss.MixLoss` (loss.py:53-86) and `loss.DINOLoss` (loss.py:89-152): same constructors,
buffers and forward signatures.  MixLoss' log-softmax / scatter / interpolation / KL chain and its backward run as one
fused row kernel (vtb_mix_loss); DINOLoss' 18 log-softmax / multiply / sum chains run as one fused kernel
(vtb_dino_loss), the centre update keeps the reference's all-reduce + EMA."""
import torch
import torch.distributed as dist
from torch import nn

from vtb200 import multi
from vtb200.blocks import DINOLossFn


class _MixLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target1, target2, interpolation, eps, reduction):
        B = output.shape[0]
        rows = reduction == "none"
        loss, row_loss, dlogits, _ = multi.mix_loss(
            output, target1, target2, interpolation, eps=eps, loss_scale=1.0 / B if reduction == "mean" else 1.0,
            want_loss=not rows, want_rows=rows, want_grad=ctx.needs_input_grad[0])
        if rows:  # the stashed gradient is d(sum of rows): scaled per row by the incoming gradient in backward
            ctx.rows = True
            ctx.save_for_backward(dlogits)
            return row_loss
        ctx.rows = False
        ctx.save_for_backward(dlogits)
        return loss[0]

    @staticmethod
    def backward(ctx, grad):
        (dlogits,) = ctx.saved_tensors
        g = dlogits * (grad.unsqueeze(-1) if ctx.rows else grad)
        return g, None, None, None, None, None


def cross_entropy(output, target):
    """nn.CrossEntropyLoss()(output, target) (train.py:155, the validation criterion) through the same row kernel:
    MixLoss with eps = 0 and no mixing partner is the cross entropy against a one-hot target.
    Labels must lie in [0, n_class): an out-of-range label (nn.CrossEntropyLoss's ignore_index = -100 is not supported;
    no reference loader produces one) makes its row contribute nothing, and the mean still divides by the full batch."""
    logits = output if output.dtype == torch.float32 else output.float()
    if logits.stride(-1) != 1:
        logits = logits.contiguous()
    return _MixLossFn.apply(logits, target.contiguous(), None, None, 0.0, "mean")


class MixLoss(nn.Module):
    def __init__(self, eps=0, reduction="mean"):
        super().__init__()

        self.eps = eps
        self.reduction = reduction

    def forward(self, output, target1, target2, interpolation):
        inter = torch.as_tensor(interpolation, dtype=torch.float32, device=output.device)
        if inter.dim() == 0:
            inter = inter.expand(output.shape[0])
        logits = output if output.dtype == torch.float32 else output.float()
        if logits.stride(-1) != 1:
            logits = logits.contiguous()
        return _MixLossFn.apply(logits, target1.contiguous(), target2.contiguous(), inter.contiguous(), self.eps,
                                self.reduction)


class DINOLoss(nn.Module):
    def __init__(
        self,
        out_dim,
        n_crop,
        warmup_teacher_temperature,
        teacher_temperature,
        warmup_teacher_epoch,
        n_epoch,
        student_temperature=0.1,
        center_momentum=0.9,
    ):
        super().__init__()

        self.student_temperature = student_temperature
        self.center_momentum = center_momentum
        self.n_crop = n_crop
        self.register_buffer("center", torch.zeros(1, out_dim))

        self.teacher_temperature_schedule = torch.cat(
            (
                torch.linspace(warmup_teacher_temperature, teacher_temperature, warmup_teacher_epoch),
                torch.ones(n_epoch - warmup_teacher_epoch) * teacher_temperature,
            )
        ).tolist()

    def forward(self, student_output, teacher_output, epoch):
        temperature = self.teacher_temperature_schedule[epoch]
        total_loss = DINOLossFn.apply(student_output.float(), teacher_output.float(), self.center, self.n_crop,
                                      self.student_temperature, temperature)
        self.update_center(teacher_output)

        return total_loss

    @torch.no_grad()
    def update_center(self, teacher_out):
        batch_center = torch.sum(teacher_out, dim=0, keepdim=True)
        world = 1
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(batch_center)
            world = dist.get_world_size()
        batch_center = batch_center / (len(teacher_out) * world)

        self.center.mul_(self.center_momentum).add_(batch_center, alpha=1 - self.center_momentum)
