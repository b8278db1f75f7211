/**
 * Self-test of the Node package, shaped after the reference's src/test.ts (same seven
 * rate / channel / quality cases, same duration assertion `|in - out| < 0.01 s`), on synthetic
 * PCM instead of the reference's resources/*.pcm so that it runs from this package alone; plus
 * the two things the reference cannot test: processChunks == per-stream processChunk, and the
 * batched Transform. Run with `npm test` on a machine with a B200 and Node >= 12.17.
 */
import SpeexResampler, { SpeexResamplerBatchTransform, SpeexResamplerTransform } from './index';

const assert = (condition: boolean, message: string) => {
  if (!condition) {
    throw new Error(message);
  }
};

const cases = [
  { inRate: 24000, outRate: 48000, channels: 1, quality: 5 },
  { inRate: 24000, outRate: 24000, channels: 2, quality: 5 },
  { inRate: 24000, outRate: 48000, channels: 2, quality: 10 },
  { inRate: 44100, outRate: 48000, channels: 2, quality: undefined },
  { inRate: 44100, outRate: 48000, channels: 2, quality: 10 },
  { inRate: 44100, outRate: 48000, channels: 2, quality: 1 },
  { inRate: 44100, outRate: 24000, channels: 2, quality: 5 },
];

/** `seconds` of interleaved int16: two sines and a little noise per channel */
function synth(channels: number, rate: number, seconds: number, seed: number): Buffer {
  const frames = Math.floor(rate * seconds);
  const buf = Buffer.alloc(frames * channels * 2);
  let state = seed >>> 0;
  for (let f = 0; f < frames; f++) {
    for (let c = 0; c < channels; c++) {
      state = (Math.imul(state, 1664525) + 1013904223) >>> 0;
      const noise = (state >>> 16) / 65536 * 4000 - 2000;
      const t = f / rate;
      const v = 8000 * Math.sin(2 * Math.PI * (220 + 37 * seed) * t) + 4000 * Math.sin(2 * Math.PI * 3.1 * 220 * t + c) + noise;
      buf.writeInt16LE(Math.max(-32768, Math.min(32767, Math.round(v))), (f * channels + c) * 2);
    }
  }
  return buf;
}

const duration = (bytes: number, rate: number, channels: number) => bytes / rate / 2 / channels;

async function oneShot() {
  for (const c of cases) {
    const r = new SpeexResampler(c.channels, c.inRate, c.outRate, c.quality);
    const pcm = synth(c.channels, c.inRate, 5, 1);
    const res = r.processChunk(pcm);
    const din = duration(pcm.length, c.inRate, c.channels);
    const dout = duration(res.length, c.outRate, c.channels);
    assert(Math.abs(din - dout) < 0.01, `Stream duration not matching target, in: ${din}s != out:${dout}`);
  }
}

async function transformStream() {
  for (const c of cases) {
    const t = new SpeexResamplerTransform(c.channels, c.inRate, c.outRate, c.quality);
    const pcm = synth(c.channels, c.inRate, 5, 2);
    let res = Buffer.alloc(0);
    t.on('data', (d: Buffer) => { res = Buffer.concat([res, d]); });
    const done = new Promise((resolve) => t.on('end', resolve));
    // 64 KiB reads like fs.createReadStream, with an odd size thrown in for the alignment carry
    const sizes = [65536, 65536, 4097, 3, 65536, 1, 30001];
    for (let pos = 0, k = 0; pos < pcm.length; k++) {
      const n = sizes[k % sizes.length];
      t.write(pcm.slice(pos, pos + n));
      pos += n;
    }
    t.end();
    await done;
    const din = duration(pcm.length, c.inRate, c.channels);
    const dout = duration(res.length, c.outRate, c.channels);
    assert(Math.abs(din - dout) < 0.01, `Stream duration not matching target, in: ${din}s != out:${dout}`);
  }
}

async function batched() {
  const S = 64, channels = 2, inRate = 44100, outRate = 48000;
  const single: SpeexResampler[] = [];
  const batch: SpeexResampler[] = [];
  for (let s = 0; s < S; s++) {
    single.push(new SpeexResampler(channels, inRate, outRate, 7));
    batch.push(new SpeexResampler(channels, inRate, outRate, 7));
  }
  const pcm = single.map((_, s) => synth(channels, inRate, 0.2, 10 + s));
  const hop = 882 * channels * 2;   // 20 ms
  for (let k = 0; k < 10; k++) {
    const chunks = pcm.map((p) => p.slice(k * hop, (k + 1) * hop));
    const got = SpeexResampler.processChunks(batch, chunks);
    chunks.forEach((c, s) => {
      const want = single[s].processChunk(c);
      assert(want.length === got[s].length, `stream ${s} hop ${k}: length ${got[s].length} != ${want.length}`);
      // the batch takes the tensor-core kernel, the single stream may take another: <= 1 LSB
      for (let i = 0; i < want.length; i += 2) {
        assert(Math.abs(want.readInt16LE(i) - got[s].readInt16LE(i)) <= 1, `stream ${s} hop ${k} sample ${i / 2}`);
      }
    });
  }
  const t = new SpeexResamplerBatchTransform(4, channels, inRate, outRate, 7);
  const outs: Buffer[][] = [];
  t.on('data', (d: Buffer[]) => outs.push(d));
  t.write([pcm[0].slice(0, 4097), pcm[1].slice(0, 3), Buffer.alloc(0), pcm[3].slice(0, hop)]);
  t.write([pcm[0].slice(4097, 8000), pcm[1].slice(3, 4000), pcm[2].slice(0, hop), pcm[3].slice(hop, 2 * hop)]);
  t.end();
  await new Promise((resolve) => t.on('end', resolve));
  assert(outs.length === 2 && outs[0].length === 4, 'batch transform emits one array per write');
}

SpeexResampler.initPromise
  .then(oneShot)
  .then(transformStream)
  .then(batched)
  .then(() => console.log('all tests passed'))
  .catch((e) => {
    console.error(e);
    process.exit(1);
  });
