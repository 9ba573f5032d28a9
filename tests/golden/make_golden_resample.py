"""Golden vectors for the ingest path: outputs of torchaudio.transforms.Resample (the library call the reference makes,
audiotoken/utils.py:41-42, 98-99) and of the reference's own convert_audio (loaded from /root/reference by AST, the
package itself cannot be imported offline).  Run in the build container:  python tests/golden/make_golden_resample.py"""
import ast, os
import numpy as np, torch, torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
src = open('/root/reference/audiotoken/utils.py').read()
tree = ast.parse(src)
fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'convert_audio')
ns = {'torch': torch, 'torchaudio': torchaudio, 'logger': type('L', (), {'warning': staticmethod(lambda *a, **k: None)})()}
exec(compile(ast.Module(body=[fn], type_ignores=[]), 'utils.py', 'exec'), ns)
convert_audio = ns['convert_audio']

save = {}
cases = [(44100, 16000, 1, 4410 * 2 + 37), (48000, 24000, 2, 9601), (8000, 16000, 1, 3000), (22050, 24000, 2, 5000),
         (16000, 16000, 2, 1234), (44100, 24000, 1, 7001), (32000, 16000, 1, 6400)]
for k, (sr, tgt, ch, n) in enumerate(cases):
    g = torch.Generator().manual_seed(100 + k)
    x = (torch.rand(ch, n, generator=g) * 2 - 1) * 0.8
    y = convert_audio(x, sr, tgt)
    save[f'case{k}_meta'] = np.array([sr, tgt, ch, n])
    save[f'case{k}_in'] = x.numpy()
    save[f'case{k}_out'] = y.numpy()
    print(k, sr, tgt, ch, n, tuple(y.shape))
np.savez_compressed(os.path.join(HERE, 'resample.npz'), **save)
