/**
 * TypeScript surface of node-speex-resampler over libspeexb200.so (B200, sm_100a).
 *
 * Same public API as the reference's src/index.ts (class SpeexResampler with `initPromise`,
 * the four public fields, synchronous `processChunk(Buffer): Buffer`; named export
 * `SpeexResamplerTransform`; default export SpeexResampler), so user code and the reference's
 * own src/test.ts run unchanged. What changed underneath: the Emscripten module and the
 * malloc/HEAPU8 staging of src/index.ts:59-115 are gone; `../build/Release/speexb200.node`
 * (src/addon.c) calls the same five C symbols natively and the arithmetic runs on the GPU.
 *
 * New: `SpeexResampler.processChunks(resamplers, chunks)` resamples many independent streams
 * in one launch (streams sharing channels / rates / quality are batched; per-stream results are
 * identical to calling `processChunk` on each), and `SpeexResamplerBatchTransform` does the same
 * for object-mode streams of chunk arrays.
 */
import { Transform, TransformCallback } from 'stream';

interface Addon {
  deviceCount(): number;
  lastError(): string;
  init(channels: number, inRate: number, outRate: number, quality: number): object;
  process(handle: object, chunk: Buffer, capFrames: number): Buffer;
  destroy(handle: object): void;
  batchCreate(nStreams: number, channels: number, inRate: number, outRate: number, quality: number, device: number): object;
  batchProcess(batch: object, chunks: Buffer[], capFrames: Uint32Array): Buffer[];
  batchAdopt(batch: object, streamIndex: number, handle: object): void;
  batchDestroy(batch: object): void;
  setKernel(handle: object, kernel: number): void;
  batchSetKernel(batch: object, kernel: number): void;
}

/** kernel families of include/speexb200.h (SPXB_KERNEL_*) */
export const KERNEL_AUTO = 0, KERNEL_STRICT = 1, KERNEL_TILED = 2, KERNEL_TENSOR = 3;

let addon: Addon | undefined;
// The reference resolves this promise when the WASM module has compiled (src/index.ts:19);
// here it resolves when the native addon is loaded and a CUDA device answers.
const globalModulePromise: Promise<any> = Promise.resolve().then(() => {
  // eslint-disable-next-line @typescript-eslint/no-var-requires
  const a: Addon = require('../build/Release/speexb200.node');
  if (a.deviceCount() <= 0) {
    throw new Error('libspeexb200: no CUDA device (' + a.lastError() + ')');
  }
  addon = a;
  return a;
});

const BYTES_PER_SAMPLE = Uint16Array.BYTES_PER_ELEMENT;

class SpeexResampler {
  _handle: object | undefined;
  // grow-only output capacity in BYTES, exactly the reference's _outBufferSize (src/index.ts:80-87)
  _outBufferSize = -1;
  // set while the stream lives inside a batch created by processChunks
  _batch: StreamGroup | undefined;
  _batchIndex = -1;

  static initPromise = globalModulePromise as Promise<any>;

  /** Kernel family (not in the reference). processChunk / the Transform default to the bit-exact
    * kernel -- the reference's bytes --, the tensor-core kernel (+-1 LSB, >= 90 dB) is opt-in there;
    * processChunks, whose purpose is throughput, defaults to AUTO (tensor when the call qualifies). */
  kernel = KERNEL_STRICT;
  static batchKernel = KERNEL_AUTO;

  /**
    * channels: interleaved channel count (>= 1); inRate / outRate: sample rates in Hz;
    * quality: Speex quality 0..10 (7 when omitted; higher = longer filter).
    * Nothing is allocated until the first chunk arrives.
    */
  constructor(
    public channels,
    public inRate,
    public outRate,
    public quality = 7) {}

  /** output capacity in frames for a chunk of `bytes` bytes: the running maximum of
    * ceil(bytes * outRate / inRate), truncated to whole frames by the 'i32' store of
    * src/index.ts:95 */
  _capacityFrames(bytes: number): number {
    const target = Math.ceil(bytes * this.outRate / this.inRate);
    if (this._outBufferSize < target) {
      this._outBufferSize = target;
    }
    return Math.trunc(this._outBufferSize / this.channels / BYTES_PER_SAMPLE);
  }

  _check(chunk: Buffer) {
    if (!addon) {
      throw new Error('You need to wait for SpeexResampler.initPromise before calling this method');
    }
    if (chunk.length % (this.channels * BYTES_PER_SAMPLE) !== 0) {
      throw new Error('Chunk length should be a multiple of channels * 2 bytes');
    }
  }

  /** One call = one chunk of interleaved little-endian int16 PCM in, the resampled chunk out. */
  processChunk(chunk: Buffer): Buffer {
    this._check(chunk);
    if (this._batch) {
      // the stream's state lives in a batch: run the whole batch with empty chunks for the others
      return this._batch.processOne(this._batchIndex, chunk);
    }
    if (!this._handle) {
      // lazy init; a failed init throws strerror(err) and leaves _handle unset so that the next
      // call retries (src/index.ts:59-65)
      this._handle = addon!.init(this.channels, this.inRate, this.outRate, this.quality);
      addon!.setKernel(this._handle, this.kernel);
    }
    return addon!.process(this._handle, chunk, this._capacityFrames(chunk.length));
  }

  /**
    * Resample one chunk of each of many independent streams in as few launches as possible.
    * resamplers[i] receives chunks[i]; the result equals resamplers.map((r, i) => r.processChunk(chunks[i])).
    * Streams with equal (channels, inRate, outRate, quality) share one device batch; the first
    * call moves each stream's state into it.
    */
  static processChunks(resamplers: SpeexResampler[], chunks: Buffer[]): Buffer[] {
    if (resamplers.length !== chunks.length) {
      throw new Error('processChunks needs one chunk per resampler');
    }
    resamplers.forEach((r, i) => r._check(chunks[i]));
    const out: Buffer[] = new Array(chunks.length);
    // group by configuration, keeping the caller's order inside each group
    const groups = new Map<string, number[]>();
    resamplers.forEach((r, i) => {
      const key = [r.channels, r.inRate, r.outRate, r.quality].join('/');
      const g = groups.get(key);
      if (g) { g.push(i); } else { groups.set(key, [i]); }
    });
    for (const idx of groups.values()) {
      const members = idx.map((i) => resamplers[i]);
      const group = StreamGroup.covering(members);
      const res = group.process(members, idx.map((i) => chunks[i]));
      idx.forEach((i, k) => { out[i] = res[k]; });
    }
    return out;
  }

  /** release the device state now instead of at garbage collection (the reference has no
    * equivalent: it never frees, src/index.ts has no destroy) */
  destroy() {
    if (this._handle && addon) {
      addon.destroy(this._handle);
    }
    this._handle = undefined;
  }
}

/** the device batch behind a set of SpeexResampler instances of one configuration */
class StreamGroup {
  batch: object;
  members: SpeexResampler[];
  caps: Uint32Array;

  constructor(members: SpeexResampler[]) {
    const r = members[0];
    this.members = members;
    this.batch = addon!.batchCreate(members.length, r.channels, r.inRate, r.outRate, r.quality, 0);
    this.caps = new Uint32Array(members.length);
    addon!.batchSetKernel(this.batch, SpeexResampler.batchKernel);
    members.forEach((m, i) => {
      if (m._handle) {
        // the stream already ran through processChunk: carry last_sample / samp_frac_num / history over
        addon!.batchAdopt(this.batch, i, m._handle);
        m.destroy();
      } else if (m._batch) {
        throw new Error('processChunks: a resampler cannot move between batches');
      }
      m._batch = this;
      m._batchIndex = i;
    });
  }

  /** the group that holds exactly these members in this order (created on first use) */
  static covering(members: SpeexResampler[]): StreamGroup {
    const g = members[0]._batch;
    if (g && g.members.length === members.length && g.members.every((m, i) => m === members[i])) {
      return g;
    }
    if (members.some((m) => m._batch)) {
      throw new Error('processChunks: call it with the same resamplers, in the same order, every time');
    }
    return new StreamGroup(members);
  }

  process(members: SpeexResampler[], chunks: Buffer[]): Buffer[] {
    members.forEach((m, i) => { this.caps[i] = m._capacityFrames(chunks[i].length); });
    return addon!.batchProcess(this.batch, chunks, this.caps);
  }

  /** one stream advances, the others get an empty chunk (their state does not move) */
  processOne(index: number, chunk: Buffer): Buffer {
    const chunks = this.members.map((_, i) => (i === index ? chunk : EMPTY_BUFFER));
    this.caps.fill(0);
    this.caps[index] = this.members[index]._capacityFrames(chunk.length);
    return addon!.batchProcess(this.batch, chunks, this.caps)[index];
  }
}

const EMPTY_BUFFER = Buffer.alloc(0);

/** bytes of `chunk` (after `carry`) that do not fill a whole frame are kept for the next chunk:
  * the alignment rule of the reference's _transform (src/index.ts:139-154) */
function alignChunk(carry: Buffer, chunk: Buffer, frameBytes: number): { whole: Buffer, rest: Buffer } {
  const joined = carry.length > 0 ? Buffer.concat([carry, chunk]) : chunk;
  const extra = joined.length % frameBytes;
  if (extra === 0) {
    return { whole: joined, rest: EMPTY_BUFFER };
  }
  return { whole: joined.slice(0, joined.length - extra), rest: Buffer.from(joined.slice(joined.length - extra)) };
}

export class SpeexResamplerTransform extends Transform {
  resampler: SpeexResampler;
  _alignementBuffer: Buffer;

  /**
    * channels: interleaved channel count (>= 1); inRate / outRate: sample rates in Hz;
    * quality: Speex quality 0..10 (7 when omitted; higher = longer filter).
    * Nothing is allocated until the first chunk arrives.
    */
  constructor(public channels, public inRate, public outRate, public quality = 7) {
    super();
    this.resampler = new SpeexResampler(channels, inRate, outRate, quality);
    this._alignementBuffer = EMPTY_BUFFER;
  }

  _transform(chunk: Buffer, encoding: BufferEncoding, callback: TransformCallback) {
    const { whole, rest } = alignChunk(this._alignementBuffer, chunk, this.channels * BYTES_PER_SAMPLE);
    this._alignementBuffer = rest;
    try {
      callback(null, this.resampler.processChunk(whole));
    } catch (e) {
      callback(e as Error);
    }
  }
}

/**
 * Object-mode Transform over MANY streams: every written object is an array with one Buffer per
 * stream (any lengths, possibly empty), every emitted object the array of resampled Buffers.
 * Each stream keeps its own alignment carry; all streams of a write go to the GPU in one launch.
 */
export class SpeexResamplerBatchTransform extends Transform {
  resamplers: SpeexResampler[];
  _carry: Buffer[];

  constructor(public streams: number, public channels, public inRate, public outRate, public quality = 7) {
    super({ objectMode: true });
    this.resamplers = [];
    this._carry = [];
    for (let i = 0; i < streams; i++) {
      this.resamplers.push(new SpeexResampler(channels, inRate, outRate, quality));
      this._carry.push(EMPTY_BUFFER);
    }
  }

  _transform(chunks: Buffer[], encoding: BufferEncoding, callback: TransformCallback) {
    if (!Array.isArray(chunks) || chunks.length !== this.streams) {
      callback(new Error('SpeexResamplerBatchTransform expects an array with one Buffer per stream'));
      return;
    }
    const frameBytes = this.channels * BYTES_PER_SAMPLE;
    const whole = chunks.map((c, i) => {
      const a = alignChunk(this._carry[i], c, frameBytes);
      this._carry[i] = a.rest;
      return a.whole;
    });
    try {
      callback(null, SpeexResampler.processChunks(this.resamplers, whole));
    } catch (e) {
      callback(e as Error);
    }
  }
}

export default SpeexResampler;
