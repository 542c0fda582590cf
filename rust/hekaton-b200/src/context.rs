//! Device contexts.  A context is one GPU + one CUDA stream + scratch memory and is NOT thread-safe: use one per
//! host thread.  Forked contexts share every uploaded resource (base sets, matrices, proving keys, NTT tables).
use crate::error::{check, HkError, Result};
use core::ptr;
use hekaton_b200_sys as sys;
use std::cell::RefCell;

pub struct Gpu {
    pub(crate) raw: *mut sys::hk_ctx,
    pub(crate) curve: u32,
    pub(crate) device: i32,
}

// A context may move between threads (it owns its stream); it must not be used from two at once (`!Sync`).
unsafe impl Send for Gpu {}

impl Gpu {
    /// `curve`: `sys::HK_CURVE_BLS12_381` or `sys::HK_CURVE_BN254` (see [`crate::GpuPairing::CURVE_ID`]).
    pub fn new(device: i32, curve: u32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::hk_ctx_create(device, curve, ptr::null_mut(), &mut raw) };
        if rc != sys::HK_OK {
            return Err(HkError {
                code: rc,
                message: "hk_ctx_create failed (an sm_100 GPU is required; there is no CPU fallback)".into(),
            });
        }
        Ok(Gpu { raw, curve, device })
    }

    /// A context on the same GPU with its own stream and scratch that sees the parent's handles.
    pub fn fork(&self) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(self.raw, unsafe { sys::hk_ctx_fork(self.raw, ptr::null_mut(), &mut raw) })?;
        Ok(Gpu { raw, curve: self.curve, device: self.device })
    }

    /// Wait for everything enqueued on this context (`HK_DEVICE` / `HK_HOST_PINNED_ASYNC` calls only enqueue).
    pub fn sync(&self) -> Result<()> {
        check(self.raw, unsafe { sys::hk_ctx_sync(self.raw) })
    }

    pub fn set_option(&self, option: u32, value: u64) -> Result<()> {
        check(self.raw, unsafe { sys::hk_ctx_set_option(self.raw, option, value) })
    }

    pub fn kernel_launches(&self) -> u64 {
        unsafe { sys::hk_launch_count(self.raw) }
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe {
            sys::hk_ctx_destroy(self.raw);
        }
    }
}

thread_local! {
    static ROOTS: RefCell<Vec<Gpu>> = RefCell::new(Vec::new());
}
static PARENTS: std::sync::Mutex<Vec<(u32, usize)>> = std::sync::Mutex::new(Vec::new());

/// The calling thread's context for `curve` on the device named by `HEKATON_B200_DEVICE` (default 0), created on
/// first use.  The first thread creates the root context; later threads fork it, so base sets and keys uploaded by
/// one thread are usable from all (the reference's worker threads share one `ProvingKey`).
pub fn with_gpu<T>(curve: u32, f: impl FnOnce(&Gpu) -> T) -> T {
    ROOTS.with(|slot| {
        let mut v = slot.borrow_mut();
        if let Some(g) = v.iter().find(|g| g.curve == curve) {
            return f(g);
        }
        let device = std::env::var("HEKATON_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut parents = PARENTS.lock().unwrap();
        let g = match parents.iter().find(|(c, _)| *c == curve) {
            Some(&(_, root)) => {
                let root = root as *mut sys::hk_ctx;
                let mut raw = ptr::null_mut();
                let rc = unsafe { sys::hk_ctx_fork(root, ptr::null_mut(), &mut raw) };
                assert_eq!(rc, sys::HK_OK, "hk_ctx_fork failed");
                Gpu { raw, curve, device }
            }
            None => {
                // the root context is intentionally leaked: it owns the shared resources for the process lifetime
                let root = Gpu::new(device, curve).expect("hekaton_b200: no usable B200 (there is no CPU fallback)");
                parents.push((curve, root.raw as usize));
                let forked = root.fork().expect("hk_ctx_fork failed");
                core::mem::forget(root);
                forked
            }
        };
        v.push(g);
        f(v.last().unwrap())
    })
}
