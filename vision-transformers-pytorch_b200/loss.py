"""Drop-in for the reference's `loss.DINOLoss` (loss.py:89-152): same constructor, buffers and forward signature; the
18 log-softmax / multiply / sum chains of the reference run as one fused kernel (vtb_dino_loss), the centre update keeps
the reference's all-reduce + EMA."""
import torch
import torch.distributed as dist
from torch import nn

from vtb200.blocks import DINOLossFn


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
