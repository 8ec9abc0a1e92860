// Type declarations of the Node-API addon built from fheb_addon.cc (load it as `require("./build/Release/node-fhe-accelerate.node")`).
//
// Part 1 repeats, member for member, the surface the reference's napi-rs addon exposes today (its generated typings:
// initialize / detectHardware / version / class ModularArithmetic), so existing callers keep compiling.  Part 2 is
// the bulk surface FHEEngineImpl needs (src/api/fhe-engine.ts:209-321).  Polynomial data are BigUint64Array,
// row-major [batch][N]; 64-bit scalars that do not fit a JS number are bigint.

declare namespace fheb {
  // ---- part 1: the existing surface --------------------------------------------------------------------
  interface HardwareCapabilities {
    hasSme: boolean; hasMetal: boolean; hasNeon: boolean; hasAmx: boolean  // always false on this backend
    metalGpuCores: number       // streaming multiprocessors (148 on a B200)
    unifiedMemorySize: number   // bytes of HBM
  }
  function initialize(): void
  function detectHardware(): HardwareCapabilities
  function version(): string
  class ModularArithmetic {
    constructor(modulus: number)
    montgomeryMul(a: number, b: number): number
    modAdd(a: number, b: number): number
    modSub(a: number, b: number): number
    toMontgomery(a: number): number
    fromMontgomery(a: number): number
    getModulus(): number
  }

  // ---- part 2: bulk entry points ---------------------------------------------------------------------------
  class NttProcessor {
    constructor(degree: number, modulus: bigint)
    forwardBatch(coeffs: BigUint64Array): BigUint64Array   // in place; returns its argument
    inverseBatch(coeffs: BigUint64Array): BigUint64Array
    polymulBatch(a: BigUint64Array, b: BigUint64Array): BigUint64Array   // PolynomialRing::multiply over a batch
  }
  class BootstrapEngine {
    // bsk: [lweDimension][(k+1)*level][k+1][degree] coefficient-form words
    constructor(degree: number, modulus: bigint, lweDimension: number, glweDimension: number, baseLog: number, level: number,
                bsk: BigUint64Array)
    bootstrapBatch(lwe: BigUint64Array, testPoly: BigUint64Array): BigUint64Array   // [batch][n+1] -> [batch][k*degree+1]
  }
  // PolynomialRing::add / subtract over whole ciphertexts or batches (element-wise, any length)
  function modAddBatch(a: BigUint64Array, b: BigUint64Array, modulus: bigint): BigUint64Array
  function modSubBatch(a: BigUint64Array, b: BigUint64Array, modulus: bigint): BigUint64Array
  // one Node process, every visible GPU: host batches of every bulk call below are spread over them; returns the count
  function setDevices(): number
  // ballots: [count][2][degree] -> [2][degree]
  function tallyVotes(ballots: BigUint64Array, degree: number, modulus: bigint): BigUint64Array
  // FHEV records back to back; status[i]: 0 ok, 1 too small, 2 bad magic, 3 checksum, 4 shape
  function ingestBallots(wire: Buffer, count: number, numChoices: number, degree: number, modulus: bigint):
    { ciphertexts: BigUint64Array, status: Uint8Array, accepted: number }
  // EncryptionEngine::multiply + relinearize; evalKeyWire: an FHEE container
  function multiplyRelinearize(degree: number, modulus: bigint, evalKeyWire: Buffer, ct1: BigUint64Array, ct2: BigUint64Array): BigUint64Array
}
export = fheb
