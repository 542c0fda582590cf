//! Error mapping: every C entry point returns 0 or a negative `HK_ERR_*` code plus a per-context message.
use core::fmt;
use hekaton_b200_sys as sys;

#[derive(Debug, Clone, PartialEq, Eq)]
pub struct HkError {
    pub code: i32,
    pub message: String,
}

impl HkError {
    /// `HK_ERR_LENGTH`: what `VariableBaseMSM::msm` reports as `Err(min_len)` (`committer.rs:89`).
    pub fn is_length_mismatch(&self) -> bool {
        self.code == sys::HK_ERR_LENGTH
    }
}

impl fmt::Display for HkError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "hekaton_b200 error {}: {}", self.code, self.message)
    }
}

impl std::error::Error for HkError {}

pub type Result<T> = core::result::Result<T, HkError>;

/// Turn a return code into a `Result`, fetching the context's message on failure.
pub(crate) fn check(ctx: *mut sys::hk_ctx, rc: i32) -> Result<()> {
    if rc == sys::HK_OK {
        return Ok(());
    }
    let message = unsafe {
        let p = sys::hk_last_error_string(ctx);
        if p.is_null() {
            String::new()
        } else {
            std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    };
    Err(HkError { code: rc, message })
}
