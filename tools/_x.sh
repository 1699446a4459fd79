cat > /tmp/r.py <<'P'
import sys; sys.path.insert(0,'.')
import sdfkit_b200 as sk
from sdfkit_b200 import scenes
sdf = scenes.readme_scene()[0].ToSdf()
for _ in range(3): img = sdf.ToImage(1920, 1080, *scenes.CAMERA)
P
ncu --set full --import-source on --clock-control none -k regex:sdfk_k_render$ --launch-skip 2 -c 1 -o gpurun_out/s3_render -f python /tmp/r.py > /dev/null 2>&1
ls -la gpurun_out/s3_render.ncu-rep
