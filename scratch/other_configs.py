"""BASELINE configs 3-5 on one B200 (fast mode): smartaverage throughput, 1080p stylisation, 1024^2 training step."""
import sys, time, json, torch
sys.path.insert(0, '.')
import artist_style_transfer_b200 as ast
dev = torch.device('cuda')
torch.manual_seed(2)
net = ast.StyleTransfer(device=dev, precision='fast'); vgg = ast.VGG16(vgg_path=None, precision='fast').to(dev)
def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
# config 3: smartaverage, 512^2 paintings (reference semantics: feature sums -> one Gram)
paint = [torch.randint(0, 256, (3, 512, 512), device=dev).float() for _ in range(32)]
ms = timed(lambda: ast.style_grams_smartaverage(vgg, paint, 1, mode="reference"), 2)
out["config3_smartaverage_paintings_per_s"] = 32 / (ms / 1e3)
ms = timed(lambda: ast.style_grams_smartaverage(vgg, paint, 1, mode="mean_gram"), 2)
out["config3_mean_gram_paintings_per_s"] = 32 / (ms / 1e3)
del paint
# config 4: 1080p stylisation, batch 8, forward only
x = torch.randint(0, 256, (8, 3, 1080, 1920), device=dev).float()
with torch.no_grad():
    y = net(x)
    assert y.shape == x.shape
    ms = timed(lambda: net(x), 3)
out["config4_1080p_images_per_s"] = 8 / (ms / 1e3)
del x, y
torch.cuda.empty_cache()
# config 5: training step 1024^2, batch 4
style = ast.style_grams_single(vgg, torch.randint(0, 256, (3, 1024, 1024), device=dev).float(), 4)
tr = ast.PerceptualTrainer(net, vgg, style)
xb = torch.randint(0, 256, (4, 3, 1024, 1024), device=dev).float()
ms = timed(lambda: tr.step(xb), 3)
out["config5_1024_train_images_per_s"] = 4 / (ms / 1e3)
out["max_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(out))
