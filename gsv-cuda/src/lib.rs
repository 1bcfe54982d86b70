//! gsv-cuda -- the B200 garbling / evaluation engine behind the reference's own seams.
//!
//! * [`GpuRecorder`] implements `CircuitContext` (src/circuit/circuit_context_trait.rs:12-27): the unchanged
//!   `#[component]` gadgets run against it once per verifying key and the topology is recorded + planned
//!   inside the library ([`Program::record`]).
//! * [`Session::garble`] replaces the per-instance `streaming_garbling::<AesNiHasher, AESAccumulatingHash>` loop
//!   of `Garbler::create` (src/cut_and_choose/garbler.rs:206-234) by ONE batched call per GPU;
//!   [`Session::evaluate`] does the same for `streaming_evaluation` (src/circuit/mod.rs:229-250);
//!   [`link`] is the garbler -> evaluator channel of examples/groth16_garble.rs:170-267 over NVLink.
//! * Labels cross as `S::to_bytes()` (16 big-endian bytes, src/core/s.rs:25-32).
//!
//! No Rust toolchain exists in the engine's build image: this crate is source-checked against the header by
//! tests/test_ffi_consistency.py and is compiled wherever the reference is.
pub mod ffi;

use std::{
    ffi::{CStr, CString},
    marker::PhantomData,
    os::raw::c_void,
};

use garbled_snark_verifier::{
    CircuitContext, Gate, S, WireId,
    circuit::{FromWires, WiresObject, component_key::ComponentKey, modes::ExecuteMode},
};

#[derive(Debug)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
pub type Result<T> = std::result::Result<T, Error>;

fn last_error(code: i32) -> Error {
    let message = unsafe { CStr::from_ptr(ffi::gsv_last_error()) }.to_string_lossy().into_owned();
    Error { code, message }
}
fn check(rc: i32) -> Result<()> {
    if rc == ffi::GSV_OK { Ok(()) } else { Err(last_error(rc)) }
}

fn to_u32(w: WireId) -> u32 {
    if w == WireId::UNREACHABLE { u32::MAX } else { w.0 as u32 }
}
fn from_u32(w: u32) -> WireId {
    if w == u32::MAX { WireId::UNREACHABLE } else { WireId(w as usize) }
}

/// `CircuitContext` that forwards the three builder calls to the library's recorder.  The library runs the
/// reference's two passes itself (credits pass with children opaque, then the execution pass; csrc/circuit.cpp):
/// a component body is invoked once for its credits template and once per output-liveness variant, and a wire
/// whose credits are 0 comes back as `WireId::UNREACHABLE`.
pub struct GpuRecorder {
    ctx: *mut ffi::GsvCtx,
}

struct Trampoline<'a, I, O, F> {
    f: &'a F,
    inputs: &'a I,
    _o: PhantomData<O>,
}

unsafe extern "C" fn component_trampoline<I, O, F>(
    ctx: *mut ffi::GsvCtx,
    user: *mut c_void,
    inputs: *const u32,
    n_in: u32,
    outputs: *mut u32,
    arity: u32,
) where
    I: WiresObject,
    O: FromWires,
    F: Fn(&mut GpuRecorder, &I) -> O,
{
    let t = unsafe { &*(user as *const Trampoline<I, O, F>) };
    // rebuild `I` over the callee-local wire ids, as the reference does for its metadata pass
    // (streaming_mode.rs:188-196)
    let ids = unsafe { std::slice::from_raw_parts(inputs, n_in as usize) };
    let mut it = ids.iter().map(|&w| from_u32(w));
    let local = t.inputs.clone_from(&mut || it.next().expect("input arity mismatch"));
    let mut rec = GpuRecorder { ctx };
    let out = (t.f)(&mut rec, &local).to_wires_vec();
    assert_eq!(out.len(), arity as usize, "component arity mismatch");
    let dst = unsafe { std::slice::from_raw_parts_mut(outputs, arity as usize) };
    for (d, w) in dst.iter_mut().zip(out) {
        *d = to_u32(w);
    }
}

impl CircuitContext for GpuRecorder {
    type Mode = ExecuteMode; // never evaluated on the host: gates are only recorded

    fn issue_wire(&mut self) -> WireId {
        from_u32(unsafe { ffi::gsv_ctx_issue_wire(self.ctx) })
    }

    fn add_gate(&mut self, g: Gate) {
        unsafe {
            ffi::gsv_ctx_add_gate(self.ctx, g.gate_type as i32, to_u32(g.wire_a), to_u32(g.wire_b), to_u32(g.wire_c))
        }
    }

    fn with_named_child<I: WiresObject, O: FromWires>(
        &mut self,
        key: ComponentKey,
        inputs: I,
        f: impl Fn(&mut Self, &I) -> O,
        arity: usize,
    ) -> O {
        let wires: Vec<u32> = inputs.to_wires_vec().into_iter().map(to_u32).collect();
        let mut outs = vec![0u32; arity];
        // the 8 key bytes already cover name, off-circuit parameters, arity and input length (component_key.rs:15-39)
        let key_c = CString::new(key.iter().map(|b| format!("{b:02x}")).collect::<String>()).unwrap();
        let tramp = Trampoline::<I, O, _> { f: &f, inputs: &inputs, _o: PhantomData };
        unsafe {
            ffi::gsv_ctx_component(
                self.ctx,
                key_c.as_ptr(),
                wires.as_ptr(),
                wires.len() as u32,
                arity as u32,
                component_trampoline::<I, O, _>,
                &tramp as *const _ as *mut c_void,
                outs.as_mut_ptr(),
            )
        };
        O::from_wires(&outs.into_iter().map(from_u32).collect::<Vec<_>>()).expect("component output arity")
    }
}

/// A recorded + planned circuit (once per topology, i.e. per verifying key).
pub struct Program {
    raw: *mut ffi::GsvProgram,
    pub info: ffi::GsvProgramInfo,
}
unsafe impl Send for Program {}
unsafe impl Sync for Program {}

impl Program {
    /// `CircuitBuilder::run_streaming` (src/circuit/mod.rs:253-301) against the recorder: `root` gets the input
    /// wires (ids 2..) and returns the output wires.
    pub fn record<F>(name: &str, n_inputs: usize, n_outputs: usize, root: F) -> Result<Self>
    where
        F: Fn(&mut GpuRecorder, &[WireId]) -> Vec<WireId>,
    {
        unsafe extern "C" fn root_trampoline<F: Fn(&mut GpuRecorder, &[WireId]) -> Vec<WireId>>(
            ctx: *mut ffi::GsvCtx,
            user: *mut c_void,
            inputs: *const u32,
            n_in: u32,
            outputs: *mut u32,
            arity: u32,
        ) {
            let f = unsafe { &*(user as *const F) };
            let ins: Vec<WireId> =
                unsafe { std::slice::from_raw_parts(inputs, n_in as usize) }.iter().map(|&w| from_u32(w)).collect();
            let out = f(&mut GpuRecorder { ctx }, &ins);
            assert_eq!(out.len(), arity as usize);
            let dst = unsafe { std::slice::from_raw_parts_mut(outputs, arity as usize) };
            for (d, w) in dst.iter_mut().zip(out) {
                *d = to_u32(w);
            }
        }
        let name_c = CString::new(name).unwrap();
        let raw = unsafe {
            ffi::gsv_program_record(
                name_c.as_ptr(),
                n_inputs as u32,
                n_outputs as u32,
                root_trampoline::<F>,
                &root as *const F as *mut c_void,
                std::ptr::null(),
            )
        };
        Self::wrap(raw)
    }

    /// One of the library's own named circuits (`gsv_program_build`), e.g. "groth16_verify_compressed".
    pub fn named(circuit: &str) -> Result<Self> {
        let c = CString::new(circuit).unwrap();
        Self::wrap(unsafe { ffi::gsv_program_build(c.as_ptr(), std::ptr::null()) })
    }

    /// The recorded circuit as its memoised template DAG (`gsv_program_export_templates`): the root template's index
    /// and the six arrays tmpl / gates / calls / items / call_wires / outs.
    pub fn export_templates(&self) -> Result<(u32, [Vec<u32>; 6])> {
        let mut sizes = [0u64; 6];
        let mut root = 0u32;
        let null = std::ptr::null_mut::<u32>();
        check(unsafe {
            ffi::gsv_program_export_templates(self.raw, sizes.as_mut_ptr(), &mut root, null, null, null, null, null, null)
        })?;
        let mut a: [Vec<u32>; 6] = std::array::from_fn(|k| vec![0u32; sizes[k] as usize]);
        let [t, g, c, i, w, o] = &mut a;
        check(unsafe {
            ffi::gsv_program_export_templates(
                self.raw,
                sizes.as_mut_ptr(),
                &mut root,
                t.as_mut_ptr(),
                g.as_mut_ptr(),
                c.as_mut_ptr(),
                i.as_mut_ptr(),
                w.as_mut_ptr(),
                o.as_mut_ptr(),
            )
        })?;
        Ok((root, a))
    }

    fn wrap(raw: *mut ffi::GsvProgram) -> Result<Self> {
        if raw.is_null() {
            return Err(last_error(ffi::GSV_ERR_INVALID));
        }
        let mut info: ffi::GsvProgramInfo = unsafe { std::mem::zeroed() };
        check(unsafe { ffi::gsv_program_get_info(raw, &mut info) })?;
        Ok(Self { raw, info })
    }
}
impl Drop for Program {
    fn drop(&mut self) {
        unsafe { ffi::gsv_program_destroy(self.raw) }
    }
}

/// What `streaming_garbling` returns per instance (`StreamingResult`, src/circuit/mod.rs:82-107), batched.
pub struct Garbled {
    pub delta: Vec<S>,
    pub false_label0: Vec<S>,
    pub true_label0: Vec<S>,
    pub input_label0: Vec<Vec<S>>,
    pub output_label0: Vec<Vec<S>>,
    /// `AESAccumulatingHash::finalize()` per instance (src/ciphertext_hasher.rs:23-33)
    pub ciphertext_commit: Vec<[u8; 16]>,
}

pub struct Evaluated {
    pub output_active: Vec<Vec<S>>,
    pub output_bits: Vec<Vec<bool>>,
    pub ciphertext_hash: Vec<[u8; 16]>,
}

fn labels(buf: &[u8]) -> Vec<S> {
    buf.chunks_exact(16).map(|c| S::from_bytes(c.try_into().unwrap())).collect()
}

/// Device state of a batch of instances of one program on one GPU.
pub struct Session<'p> {
    raw: *mut ffi::GsvSession,
    prog: &'p Program,
    n: usize,
}
unsafe impl Send for Session<'_> {}

impl<'p> Session<'p> {
    /// `ct_mode`: `ffi::GSV_CT_COMMIT_HOST` for verifier-scale cut-and-choose batches, `GSV_CT_COMMIT` for batches of
    /// hundreds of small circuits, `GSV_CT_KEEP` to keep the stream for `read_ciphertexts` / same-GPU evaluation.
    pub fn new(prog: &'p Program, n_instances: usize, device: i32, ct_mode: i32) -> Result<Self> {
        let opt = ffi::GsvSessionOptions {
            device,
            n_instances: n_instances as u32,
            group: 0,
            worker_threads: 0,
            ct_mode: ct_mode as u32,
            ct_ring_log2: 0,
            exec_mode: 0,
            sm_limit: 0,
            ct_buffer_bytes: 0,
            host_threads: 0,
            reserved: 0,
        };
        let raw = unsafe { ffi::gsv_session_create(prog.raw, &opt) };
        if raw.is_null() {
            return Err(last_error(ffi::GSV_ERR_CUDA));
        }
        Ok(Self { raw, prog, n: n_instances })
    }

    /// The garbling stage of `Garbler::create` for `seeds.len()` instances at once (garbler.rs:206-234).
    pub fn garble(&mut self, seeds: &[u64], hasher: i32) -> Result<Garbled> {
        assert_eq!(seeds.len(), self.n);
        let (n_in, n_out) = (self.prog.info.n_inputs as usize, self.prog.info.n_outputs as usize);
        let mut delta = vec![0u8; 16 * self.n];
        let mut fl = vec![0u8; 16 * self.n];
        let mut tl = vec![0u8; 16 * self.n];
        let mut il = vec![0u8; 16 * self.n * n_in];
        let mut ol = vec![0u8; 16 * self.n * n_out];
        let mut cc = vec![0u8; 16 * self.n];
        let mut res: ffi::GsvGarbleResult = unsafe { std::mem::zeroed() };
        res.delta = delta.as_mut_ptr();
        res.false_label0 = fl.as_mut_ptr();
        res.true_label0 = tl.as_mut_ptr();
        res.input_label0 = il.as_mut_ptr();
        res.output_label0 = ol.as_mut_ptr();
        res.ct_commit = cc.as_mut_ptr();
        check(unsafe { ffi::gsv_garble_batch(self.raw, hasher, seeds.as_ptr(), &mut res) })?;
        Ok(Garbled {
            delta: labels(&delta),
            false_label0: labels(&fl),
            true_label0: labels(&tl),
            input_label0: il.chunks_exact(16 * n_in.max(1)).map(labels).collect(),
            output_label0: ol.chunks_exact(16 * n_out.max(1)).map(labels).collect(),
            ciphertext_commit: cc.chunks_exact(16).map(|c| c.try_into().unwrap()).collect(),
        })
    }

    /// `streaming_evaluation` for the batch; `ct_streams`: one `gc_{i}.bin` image per instance, or `None` for the
    /// session's own kept stream / the linked garbler.
    pub fn evaluate(
        &mut self,
        hasher: i32,
        true_label1: &[S],
        false_label0: &[S],
        input_active: &[Vec<S>],
        input_bits: &[Vec<bool>],
        ct_streams: Option<&[&[u8]]>,
    ) -> Result<Evaluated> {
        let (n_in, n_out) = (self.prog.info.n_inputs as usize, self.prog.info.n_outputs as usize);
        let flat = |v: &[S]| v.iter().flat_map(|s| s.to_bytes()).collect::<Vec<u8>>();
        let tl = flat(true_label1);
        let fl = flat(false_label0);
        let ia: Vec<u8> = input_active.iter().flat_map(|v| flat(v)).collect();
        let ib: Vec<u8> = input_bits.iter().flat_map(|v| v.iter().map(|&b| b as u8)).collect();
        assert_eq!(ia.len(), 16 * self.n * n_in);
        let mut oa = vec![0u8; 16 * self.n * n_out];
        let mut ob = vec![0u8; self.n * n_out];
        let mut cc = vec![0u8; 16 * self.n];
        let ptrs: Vec<*const u8> = ct_streams.map(|s| s.iter().map(|x| x.as_ptr()).collect()).unwrap_or_default();
        let mut io: ffi::GsvEvaluateIo = unsafe { std::mem::zeroed() };
        io.true_label = tl.as_ptr();
        io.false_label = fl.as_ptr();
        io.input_active = ia.as_ptr();
        io.input_bits = ib.as_ptr();
        if let Some(s) = ct_streams {
            io.ct_streams = ptrs.as_ptr();
            io.ct_stream_len = (s[0].len() / 16) as u64;
        }
        io.output_active = oa.as_mut_ptr();
        io.output_bits = ob.as_mut_ptr();
        io.ct_commit = cc.as_mut_ptr();
        check(unsafe { ffi::gsv_evaluate_batch(self.raw, hasher, &mut io) })?;
        Ok(Evaluated {
            output_active: oa.chunks_exact(16 * n_out.max(1)).map(labels).collect(),
            output_bits: ob.chunks_exact(n_out.max(1)).map(|c| c.iter().map(|&b| b != 0).collect()).collect(),
            ciphertext_hash: cc.chunks_exact(16).map(|c| c.try_into().unwrap()).collect(),
        })
    }

    /// `gc_{i}.bin` bytes of one instance (needs `GSV_CT_KEEP`).
    pub fn read_ciphertexts(&mut self, instance: usize) -> Result<Vec<u8>> {
        let n = self.prog.info.n_ciphertexts;
        let mut out = vec![0u8; 16 * n as usize];
        check(unsafe { ffi::gsv_session_read_ciphertexts(self.raw, instance as u32, 0, n, out.as_mut_ptr()) })?;
        Ok(out)
    }
}
impl Drop for Session<'_> {
    fn drop(&mut self) {
        unsafe { ffi::gsv_session_destroy(self.raw) }
    }
}

/// Garbler -> evaluator streaming (examples/groth16_garble.rs:170-267): afterwards `garbler.garble(..)` and
/// `evaluator.evaluate(.., None)` are called concurrently from two threads, once per run.
pub fn link(garbler: &mut Session<'_>, evaluator: &mut Session<'_>, ring_bytes: u64) -> Result<()> {
    check(unsafe { ffi::gsv_session_link(garbler.raw, evaluator.raw, ring_bytes) })
}

/// `commit_label` for many labels (src/cut_and_choose/mod.rs:41-48).
pub fn commit_labels(device: i32, labels_in: &[S]) -> Result<Vec<[u8; 16]>> {
    let buf: Vec<u8> = labels_in.iter().flat_map(|s| s.to_bytes()).collect();
    let mut out = vec![0u8; buf.len()];
    check(unsafe { ffi::gsv_commit_labels(device, buf.as_ptr(), labels_in.len() as u64, out.as_mut_ptr()) })?;
    Ok(out.chunks_exact(16).map(|c| c.try_into().unwrap()).collect())
}
