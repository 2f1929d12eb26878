-- | The `deflate` executable of pure-zlib (reference: Deflate.hs:15-48 at the root of the checkout) over this
-- package's "Codec.Compression.Zlib": `deflate foo.z` writes `foo` by driving the incremental decoder --
-- feed a strict chunk on NeedMore, append on Chunk, stop on Done / DecompError.  Same messages, same file
-- naming rule.  UNBUILT in this repository (no GHC in the build image); the executable Python mirror that the
-- tests drive is pure_zlib_b200/deflate_cli.py.
module Main (main) where

import Codec.Compression.Zlib (ZlibDecoder (..), decompressIncremental)
import Control.Monad.ST (RealWorld, ST, stToIO)
import qualified Data.ByteString as S
import qualified Data.ByteString.Lazy as L
import Data.List (isSuffixOf)
import System.Environment (getArgs)
import System.IO (Handle, IOMode (WriteMode), hClose, openFile)

main :: IO ()
main = getArgs >>= \argv -> case argv of
  [path]
    | ".z" `isSuffixOf` path -> do
        input <- L.readFile path
        out <- openFile (take (length path - 2) path) WriteMode
        pump out (L.toChunks input) decompressIncremental
    | otherwise -> putStrLn "Unexpected file name."
  _ -> putStrLn "USAGE: deflate [filename]"

-- One decoder state at a time; the handle is closed on every terminal state (Deflate.hs:30-48).
pump :: Handle -> [S.ByteString] -> ST RealWorld (ZlibDecoder RealWorld) -> IO ()
pump out chunks step = stToIO step >>= \st -> case st of
  Chunk bytes rest -> S.hPut out bytes >> pump out chunks rest
  NeedMore feed -> case chunks of
    c : cs -> pump out cs (feed c)
    [] -> putStrLn "ERROR: Ran out of data mid-decompression." >> hClose out
  DecompError e -> putStrLn ("ERROR: " ++ show e) >> hClose out
  Done -> do
    if null chunks then return () else putStrLn "WARNING: Finished decompression with data left."
    hClose out
