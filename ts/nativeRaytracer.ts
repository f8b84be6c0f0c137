/* nativeRaytracer.ts — the body of `generateImageBuffer` (src/raytracer.ts:39-113) with the parallel renderer replaced by
 * native dispatch.  NOT COMPILED HERE (no node/tsc in the build image); Python mirror with the same structure, exercised by
 * the test-suite: mcp_raytracer_b200/raytracer.py.
 *
 *   parallel: false  -> one GPU,  camera.renderAsync(pixelData)                        (was camera.render on the main thread)
 *   parallel: true   -> all GPUs, rt_multi_* inside ONE call                           (was N worker_threads over row strips,
 *                       src/raytracer.ts:60-90, src/render-utils/renderWorker.ts:17-35; `threads` = number of GPUs)
 *   scene.camera     -> applied (the reference's MCP tool accepts and drops it, src/mcp.ts:132-142)
 *   PNG              -> sharp as before (src/raytracer.ts:102-110), on its own libvips thread: the NEXT request's render runs
 *                       on the GPU while this image is being encoded (nothing in this function holds the GPU after the
 *                       pixels are back in host memory)
 */
import sharp from 'sharp';
import { generateSceneData } from '../src/scenes/scenes.js';
import type { SceneConfig } from '../src/scenes/sceneData.js';
import { NativeCamera, applySceneCameraOptions } from './nativeCamera.js';

export async function generateImageBuffer(sceneConfig: SceneConfig & { camera?: Record<string, any> } = { type: 'default' },
                                          options: { parallel?: boolean, threads?: number, verbose?: boolean } = {}): Promise<Buffer> {
  const sceneData = applySceneCameraOptions(generateSceneData(sceneConfig), (sceneConfig as any).camera);
  const camera = new NativeCamera(sceneData, (sceneConfig as any).render, { nDevices: options.threads ?? 0 }, options.parallel ?? false);
  try {
    const pixelData = new Uint8ClampedArray(camera.imageWidth * camera.imageHeight * camera.channels);
    const stats = await camera.renderAsync(pixelData);
    if (options.verbose) {
      console.error(`Adaptive sampling stats: avg=${stats.samples.avg.toFixed(2)}, min=${stats.samples.min}, max=${stats.samples.max}`);
      console.error(`Ray bounce stats: avg=${stats.bounces.avg.toFixed(2)}, min=${stats.bounces.min}, max=${stats.bounces.max}`);
    }
    if (pixelData.length === 0) throw new Error('Generated pixelData buffer is empty before calling sharp.');   // src/raytracer.ts:97-99
    return await sharp(Buffer.from(pixelData.buffer), { raw: { width: camera.imageWidth, height: camera.imageHeight, channels: 3 } }).png().toBuffer();
  } finally {
    camera.destroy();
  }
}
