"""TEST INFRASTRUCTURE ONLY -- functional torch-fp32 restatement of the HMR image encoder + theta regressor.

Follows networks/hmr.py with the reference's ``state_dict`` (same keys):
  PreActBottleneck.forward     :93-116   (preact = relu(bn1(x)); shortcut(preact) or subsample(x, stride); 1x1, 3x3, 1x1+bias)
  PreActResNet._make_layer     :133-146  (stride on the LAST block of layers 1-3; layer4 stride 1)
  HumanModelRecovery.forward   :275-300  (conv1 7x7 s2 + bias, max_pool2d(3, 2, ceil_mode=True), layers, relu(post_bn), avg_pool2d(7))
  ThetaRegressor.forward       :236-252  (theta = mean_theta; 3 x theta += fc3(relu(fc2(relu(fc1(cat[x, theta]))))), eval: no dropout)
Pinned by tests/golden/make_hmr_golden.py, which imports the REFERENCE module (h5py / ipdb stubbed), loads the same
deterministic weights and compares full outputs; the committed tests/golden/hmr.npz carries them to the GPU box.
"""
import torch
import torch.nn.functional as F

LAYERS = ((64, 3, 2), (128, 4, 2), (256, 6, 2), (512, 3, 1))        # planes, blocks, stride of the last block


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'], sd[p + '.bias'], False, 0.0, 1e-5)


def bottleneck(x, sd, p, stride):
    preact = F.relu(_bn(x, sd, p + '.bn1'))
    if (p + '.shortcut.0.weight') in sd:
        shortcut = F.conv2d(preact, sd[p + '.shortcut.0.weight'], sd[p + '.shortcut.0.bias'], stride=stride)
    else:
        shortcut = x if stride == 1 else F.max_pool2d(x, [1, 1], stride=stride)      # subsample(), hmr.py:21-36
    o = F.relu(_bn(F.conv2d(preact, sd[p + '.conv1.weight']), sd, p + '.bn2'))
    o = F.relu(_bn(F.conv2d(o, sd[p + '.conv2.weight'], stride=stride, padding=1), sd, p + '.bn3'))
    o = F.conv2d(o, sd[p + '.conv3.weight'], sd[p + '.conv3.bias'])
    return o + shortcut


def encoder(x, sd, p='resnet'):
    out = F.conv2d(x, sd[p + '.conv1.weight'], sd[p + '.conv1.bias'], stride=2, padding=3)
    out = F.max_pool2d(out, kernel_size=3, stride=2, ceil_mode=True)
    for li, (planes, nblocks, stride) in enumerate(LAYERS):
        for bi in range(nblocks):
            out = bottleneck(out, sd, '%s.layer%d.%d' % (p, li + 1, bi), stride if bi == nblocks - 1 else 1)
    out = F.relu(_bn(out, sd, p + '.post_bn'))
    out = F.avg_pool2d(out, 7)
    return out.view(out.size(0), -1)


def regressor(feat, sd, p='regressor', iterations=3):
    theta = sd[p + '.mean_theta'].repeat(feat.shape[0], 1)
    for _ in range(iterations):
        h = torch.cat([feat, theta], dim=1)
        h = F.relu(F.linear(h, sd[p + '.fc_blocks.fc1.weight'], sd[p + '.fc_blocks.fc1.bias']))
        h = F.relu(F.linear(h, sd[p + '.fc_blocks.fc2.weight'], sd[p + '.fc_blocks.fc2.bias']))
        theta = theta + F.linear(h, sd[p + '.fc_blocks.fc3.weight'], sd[p + '.fc_blocks.fc3.bias'])
    return theta


def forward(x, sd):
    """HumanModelRecovery.forward: images [N,3,224,224] in [-1,1] -> theta [N,85]."""
    return regressor(encoder(x, sd), sd)
