"""Point-cloud tokenizer (reference: models/mla/pointcloud/backbone/{pointvit,Point_PN}.py) — placeholder module
tree with the reference's parameter names; the CUDA forward is wired in pointcloud_impl (see below)."""
from __future__ import annotations

import torch
import torch.nn as nn


class Linear1Layer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, bias=True):
        super().__init__()
        self.act = nn.ReLU(inplace=True)
        self.net = nn.Sequential(nn.Conv1d(in_channels, out_channels, kernel_size, bias=bias),
                                 nn.BatchNorm1d(out_channels), self.act)


class Linear2Layer(nn.Module):
    def __init__(self, in_channels, kernel_size=1, groups=1, bias=True, adapter_layer=0):
        super().__init__()
        self.act = nn.ReLU(inplace=True)
        mid = 32 if adapter_layer == 2 else int(in_channels / 2)
        self.net1 = nn.Sequential(nn.Conv2d(in_channels, mid, kernel_size, groups=groups, bias=bias),
                                  nn.BatchNorm2d(mid), self.act)
        self.net2 = nn.Sequential(nn.Conv2d(mid, in_channels, kernel_size, bias=bias), nn.BatchNorm2d(in_channels))


class LGA(nn.Module):
    def __init__(self, out_dim, alpha, beta, block_num, dim_expansion, type, adapter_layer=0):
        super().__init__()
        self.type, self.out_dim, self.alpha, self.beta = type, out_dim, alpha, beta
        self.linear2 = nn.Sequential(*[Linear2Layer(out_dim, bias=True, adapter_layer=adapter_layer)
                                       for _ in range(block_num)])


class EncP(nn.Module):
    def __init__(self, in_channels, input_points, num_stages, embed_dim, k_neighbors, alpha, beta, LGA_block,
                 dim_expansion, type):
        super().__init__()
        self.input_points, self.num_stages, self.embed_dim = input_points, num_stages, embed_dim
        self.alpha, self.beta, self.k_neighbors = alpha, beta, k_neighbors
        self.raw_point_embed = Linear1Layer(in_channels, embed_dim, bias=False)
        self.LGA_list = nn.ModuleList()
        out_dim, self.group_nums, self.out_dims = embed_dim, [], []
        group_num = input_points
        for i in range(num_stages):
            out_dim *= dim_expansion[i]
            group_num //= 2
            self.group_nums.append(group_num)
            self.out_dims.append(out_dim)
            self.LGA_list.append(LGA(out_dim, alpha, beta, LGA_block[i], dim_expansion[i], type, adapter_layer=i))


class Point_PN_scan(nn.Module):
    def __init__(self, in_channels=3, class_num=15, input_points=1024, num_stages=2, embed_dim=96, k_neighbors=81,
                 beta=100, alpha=1000, LGA_block=(2, 1, 1, 1), dim_expansion=(2, 2, 2, 1), type="scan"):
        super().__init__()
        self.EncP = EncP(in_channels, input_points, num_stages, embed_dim, k_neighbors, alpha, beta, list(LGA_block),
                         list(dim_expansion), type)
        self.out_channels = embed_dim
        for i in dim_expansion:
            self.out_channels *= i


class PointTokenizer(nn.Module):
    def __init__(self, in_channels=3, embed_dim=768, depth=12, num_heads=6, mlp_ratio=4., target_token_count=256,
                 norm_args=None, **kwargs):
        super().__init__()
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = Point_PN_scan()
        self.proj = nn.Linear(384, 768)
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, target_token_count + 1, embed_dim))
        self.norm = None   # create_norm(...) yields no parameters in the reference's state_dict
        self.initialize_weights()

    def initialize_weights(self):
        torch.nn.init.normal_(self.cls_token, std=.02)
        torch.nn.init.normal_(self.pos_embed, std=.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                torch.nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, (nn.LayerNorm, nn.GroupNorm, nn.BatchNorm2d, nn.BatchNorm1d)):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    def forward(self, p, x=None, **kwargs):
        from .pointcloud_impl import point_tokenizer_forward
        return point_tokenizer_forward(self, p)
